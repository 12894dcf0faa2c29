import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: minutes of CPU time (KAT pinning at 65536^2)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Built on demand."""
    from oracle import ising_oracle
    ising_oracle.lib()
    return ising_oracle


@pytest.fixture(scope="session")
def ig():
    """The product package; the shared library must have been built (make / build())."""
    import ising_gpu_b200
    ising_gpu_b200.load_library()
    return ising_gpu_b200

"""Tensor-core variant against golden outputs of the reference's tensorcore/ program run on a
B200 (tests/golden/reference_tc_b200.*, generator tools/gen_golden_tc.sh): final lattice and
the printed average magnetisation.  CPU part pins the oracle, GPU part the CUDA path."""
import io
import json
import os
import tarfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
IDX = json.load(open(os.path.join(HERE, "golden", "reference_tc_b200.json")))["cases"]


def _opts(args):
    g = lambda k, d: (args[args.index(k) + 1] if k in args else d)   # noqa: E731
    return (int(g("-x", 10)), int(g("-y", 10)), int(g("-w", 100)), int(g("-n", 1000)), float(g("-a", 0.1)),
            int(g("-s", 1234)))


def _final(case, ny, nx):
    with tarfile.open(os.path.join(HERE, "golden", "reference_tc_b200_files.tgz")) as t:
        bits = np.load(io.BytesIO(t.extractfile(case + "/final.npy").read()))
    return np.unpackbits(bits)[:ny * nx].reshape(ny, nx)


def _mag_string(total, n):
    return "%g" % abs(total / float(n))      # std::cout << double, default precision 6


@pytest.mark.parametrize("case", ["t_small", "t_rect", "t_hot", "t_mid", "t_mid_cold"])
def test_oracle_reproduces_reference_tensorcore_run(oracle, case):
    nbx, nby, w, n, alpha, seed = _opts(IDX[case]["args"])
    T = oracle.TensorcoreOracle(nbx, nby, seed=seed, alpha=alpha)
    T.init()
    T.update(w + n)
    assert _mag_string(T.sum(), T.nx * T.ny) == IDX[case]["magnetism"]
    if IDX[case]["files"]:
        assert np.array_equal(_final(case, T.ny, T.nx), (T.true_lattice() > 0).astype(np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("nn_path", [0, 1])
@pytest.mark.parametrize("case", ["t_small", "t_rect", "t_hot", "t_mid", "t_mid_cold"])
def test_gpu_variant_reproduces_reference_tensorcore_run(case, nn_path):
    from ising_gpu_b200.tensorcore import TensorcoreEngine
    nbx, nby, w, n, alpha, seed = _opts(IDX[case]["args"])
    with TensorcoreEngine(nbx, nby, seed=seed, alpha=alpha, nn_path=nn_path) as T:
        T.init()
        T.update(w + n)
        assert _mag_string(T.sum(), T.nx * T.ny) == IDX[case]["magnetism"]
        if IDX[case]["files"]:
            assert np.array_equal(_final(case, T.ny, T.nx), (T.true_lattice() > 0).astype(np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["t_big_T1.5", "t_big_Tc"])
def test_gpu_variant_full_size_magnetisation(case):
    """BASELINE.json config 4: 32768 x 32768, 100 warm-up + 1000 iterations, reference defaults."""
    from ising_gpu_b200.tensorcore import TensorcoreEngine
    nbx, nby, w, n, alpha, seed = _opts(IDX[case]["args"])
    with TensorcoreEngine(nbx, nby, seed=seed, alpha=alpha) as T:
        T.init()
        T.update(w + n)
        assert _mag_string(T.sum(), T.nx * T.ny) == IDX[case]["magnetism"]

"""ising_gpu_b200 -- B200-native (sm_100a) checkerboard Metropolis engine for the 2-D Ising
model; drop-in for the hot path of NVIDIA/ising-gpu's optimized/ program.

The product is the C-ABI library ``libising_b200.so`` (include/ising_b200.h) built from
``csrc/``; this package is the thin Python host mirror used by tests and bench.py.
There is no CPU fallback: every compute call goes to the CUDA library and raises if it
is missing or no sm_100 GPU is visible.
"""
from .engine import (CRIT_TEMP, DEFAULT_SEED, IsingEngine, IsingError, accept_count,  # noqa: F401
                     accept_table, lib_path, load_library, resolve_temperature, site_random)

__all__ = ["IsingEngine", "IsingError", "load_library", "lib_path", "accept_count", "accept_table",
           "resolve_temperature", "site_random", "DEFAULT_SEED", "CRIT_TEMP"]

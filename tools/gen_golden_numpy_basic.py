#!/usr/bin/env python
"""Golden vectors for the numpy CPU baseline (oracle/numpy_basic.py), generated FROM THE REFERENCE:
the `update_lattice` kernel of /root/reference/basic_python/ising_basic.py (lines 106-134) is exec'd
verbatim under numba's CUDA simulator (NUMBA_ENABLE_CUDASIM=1; the program itself cannot be imported
without cupy / mpi4py) on seeded int8 lattices and float32 uniforms.  Output: tests/golden/numpy_basic_golden.npz
(inputs + the lattices after every half-sweep).  Run in the build container, where /root/reference exists."""
import os
import sys

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
import math  # noqa: E402,F401  (used by the exec'd kernel)

import numpy as np  # noqa: E402
from numba import cuda  # noqa: E402

REF = os.environ.get("REF", "/root/reference") + "/basic_python/ising_basic.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_kernel(inv_temp):
    """update_lattice exactly as the reference defines it (source lines 104-134 exec'd, not copied)."""
    lines = open(REF).read().split("\n")
    a = next(i for i, l in enumerate(lines) if l.startswith("def update_lattice("))
    assert lines[a - 1].strip() == "@cuda.jit", "reference layout changed"
    b = next(i for i in range(a + 1, len(lines)) if lines[i] and not lines[i][0].isspace())
    ns = {"cuda": cuda, "math": math, "inv_temp": inv_temp}
    exec("\n".join(lines[a - 1:b]), ns)
    return ns["update_lattice"], (a, b)


def run_reference(black, white, rand, inv_temp):
    """rand: (2*sweeps, n, m/2) uniforms, black half-sweep first.  Returns the states after every half-sweep."""
    k, (a, b) = reference_kernel(inv_temp)
    n, m2 = black.shape
    threads = 128
    blocks = (n * m2 + threads - 1) // threads
    bl, wh = black.copy(), white.copy()
    states = []
    for h in range(rand.shape[0]):
        if h % 2 == 0:
            k[blocks, threads](bl, wh, rand[h], True)
        else:
            k[blocks, threads](wh, bl, rand[h], False)
        states.append((bl.copy(), wh.copy()))
    return states, (a, b)


def main():
    rng = np.random.default_rng(20261017)
    cases = {}
    for name, n, m, T, sweeps in (("a", 16, 32, 2.5, 3), ("b", 12, 20, 1.5, 2), ("c", 8, 16, 2.269185, 2)):
        black = np.where(rng.random((n, m // 2), dtype=np.float32) > 0.5, 1, -1).astype(np.int8)
        white = np.where(rng.random((n, m // 2), dtype=np.float32) > 0.5, 1, -1).astype(np.int8)
        rand = rng.random((2 * sweeps, n, m // 2), dtype=np.float32)
        states, span = run_reference(black, white, rand, 1.0 / T)
        cases[name + "_black0"] = black; cases[name + "_white0"] = white; cases[name + "_rand"] = rand
        cases[name + "_T"] = np.float64(T)
        cases[name + "_black"] = np.stack([s[0] for s in states]); cases[name + "_white"] = np.stack([s[1] for s in states])
    out = os.path.join(ROOT, "tests", "golden", "numpy_basic_golden.npz")
    np.savez_compressed(out, source="basic_python/ising_basic.py lines %d-%d under NUMBA_ENABLE_CUDASIM=1" %
                        (span[0], span[1]), **cases)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    sys.exit(main())

"""numpy restatement of basic_python/ising_basic.py's update kernel (TEST/BENCH INFRASTRUCTURE).

The reference program is numba-CUDA + cupy + mpi4py and cannot be imported without a GPU stack;
its per-site rule is restated here with whole-array numpy operations:
  update_lattice     basic_python/ising_basic.py:106-134
  generate_lattice   :69-71   (+1 if r > 0.5 else -1)
  sweep order        :175-195 (fresh uniforms, black, fresh uniforms, white)
  metric             :254     (n*m*iters / t * 1e-9 "updates per ns")
PINNED against the reference kernel itself: tools/gen_golden_numpy_basic.py exec's lines 106-134 of
the reference file under numba's CUDA simulator on seeded lattices/uniforms and stores every half-sweep
in tests/golden/numpy_basic_golden.npz; tests/test_numpy_baseline.py checks `_update` against it (and
against the live reference kernel when /root/reference is present).  What stays unpinned is only the
random stream: the reference draws its uniforms with cuRAND's host API in bulk order, not reproducible
on a CPU, so against the GPU path it is a physics (3 sigma, tests/test_gpu_physics_3sigma.py) and
throughput baseline, not a bit-level one.
"""
import time

import numpy as np

TCRIT = 2.26918531421   # basic_python/ising_basic.py:33


def _update(lattice, op, randvals, is_black, inv_temp):
    n, m = lattice.shape
    rows = np.arange(n)[:, None]
    up = np.roll(op, 1, axis=0)      # op[inn, j]
    down = np.roll(op, -1, axis=0)   # op[ipp, j]
    left = np.roll(op, 1, axis=1)    # op[i, jnn]
    right = np.roll(op, -1, axis=1)  # op[i, jpp]
    odd = (rows % 2).astype(bool)
    side = np.where(odd, right, left) if is_black else np.where(odd, left, right)
    nn = up.astype(np.int32) + down + op + side
    acc = np.exp(-2.0 * inv_temp * nn * lattice)           # float64, as math.exp in the kernel
    flip = randvals < acc
    lattice[flip] = -lattice[flip]


def run(n=256, m=256, alpha=2.5 / TCRIT, sweeps=128, seed=1234, return_state=False):
    """Returns (flips_per_ns, abs_magnetisation, seconds) [+ (black, white) with return_state]."""
    rng = np.random.default_rng(seed)
    inv_temp = 1.0 / (alpha * TCRIT)
    black = np.where(rng.random((n, m // 2), dtype=np.float32) > 0.5, 1, -1).astype(np.int8)
    white = np.where(rng.random((n, m // 2), dtype=np.float32) > 0.5, 1, -1).astype(np.int8)
    t0 = time.perf_counter()
    for _ in range(sweeps):
        _update(black, white, rng.random((n, m // 2), dtype=np.float32), True, inv_temp)
        _update(white, black, rng.random((n, m // 2), dtype=np.float32), False, inv_temp)
    t = time.perf_counter() - t0
    mag = abs(int(black.sum(dtype=np.int64)) + int(white.sum(dtype=np.int64))) / float(n * m)
    if return_state:
        return n * m * sweeps / t * 1e-9, mag, t, (black, white)
    return n * m * sweeps / t * 1e-9, mag, t


def energy_per_site(black, white):
    """-(1/N) sum s_i (s_right + s_down) for the two colour arrays (same geometry as above)."""
    n, m2 = black.shape
    lat = np.empty((n, 2 * m2), dtype=np.int32)
    even = (np.arange(n) % 2 == 0)
    lat[even, 0::2] = black[even]; lat[even, 1::2] = white[even]
    lat[~even, 0::2] = white[~even]; lat[~even, 1::2] = black[~even]
    return -float((lat * (np.roll(lat, -1, 1) + np.roll(lat, -1, 0))).sum()) / lat.size


if __name__ == "__main__":
    print("flips/ns %.4f  |m| %.4f  %.2f s" % run())

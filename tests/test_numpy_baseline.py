"""The numpy CPU baseline (oracle/numpy_basic.py, BASELINE config 1) against the reference's own kernel.

Golden vectors: tests/golden/numpy_basic_golden.npz, written by tools/gen_golden_numpy_basic.py, which exec's
/root/reference/basic_python/ising_basic.py:106-134 (`update_lattice`) verbatim under NUMBA_ENABLE_CUDASIM=1.
When the reference tree and numba are present (the build container) the same comparison also runs live."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import numpy_basic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "numpy_basic_golden.npz")


def _replay(g, name):
    black, white = g[name + "_black0"].copy(), g[name + "_white0"].copy()
    rand, inv_temp = g[name + "_rand"], 1.0 / float(g[name + "_T"])
    for h in range(rand.shape[0]):
        if h % 2 == 0:
            numpy_basic._update(black, white, rand[h], True, inv_temp)
        else:
            numpy_basic._update(white, black, rand[h], False, inv_temp)
        assert np.array_equal(black, g[name + "_black"][h]), (name, h, "black")
        assert np.array_equal(white, g[name + "_white"][h]), (name, h, "white")


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_update_equals_reference_kernel_golden(name):
    g = np.load(GOLD)
    assert "ising_basic.py" in str(g["source"])
    _replay(g, name)


def test_update_equals_reference_kernel_live(tmp_path):
    """Regenerate the vectors from /root/reference right now and compare (skipped where it cannot run)."""
    if not os.path.exists("/root/reference/basic_python/ising_basic.py"):
        pytest.skip("reference tree not present on this machine")
    pytest.importorskip("numba")
    env = dict(os.environ, NUMBA_ENABLE_CUDASIM="1")
    code = ("import sys; sys.path.insert(0, %r); import numpy as np\n"
            "import tools.gen_golden_numpy_basic as G\n"
            "rng = np.random.default_rng(7)\n"
            "n, m, T = 10, 24, 2.5\n"
            "b = np.where(rng.random((n, m // 2), dtype=np.float32) > 0.5, 1, -1).astype(np.int8)\n"
            "w = np.where(rng.random((n, m // 2), dtype=np.float32) > 0.5, 1, -1).astype(np.int8)\n"
            "r = rng.random((4, n, m // 2), dtype=np.float32)\n"
            "st, _ = G.run_reference(b, w, r, 1.0 / T)\n"
            "np.savez(%r, a_black0=b, a_white0=w, a_rand=r, a_T=np.float64(T), a_black=np.stack([s[0] for s in st]),"
            " a_white=np.stack([s[1] for s in st]))\n" % (ROOT, str(tmp_path / "live.npz")))
    subprocess.run([sys.executable, "-c", code], check=True, env=env, cwd=ROOT)
    _replay(np.load(tmp_path / "live.npz"), "a")


def test_config1_runs_and_is_disordered():
    """BASELINE config 1 (256 x 256, T = 2.5, 128 sweeps): above T_c the magnetisation stays small and the energy
    per site sits near the exact infinite-lattice value (-1.106 at T = 2.5, Onsager)."""
    fl, mag, secs, (black, white) = numpy_basic.run(n=256, m=256, sweeps=128, seed=1234, return_state=True)
    assert fl > 0 and secs > 0
    assert mag < 0.2
    e = numpy_basic.energy_per_site(black, white)
    assert -1.25 < e < -0.95, e

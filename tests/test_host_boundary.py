"""CPU tests of the product's host side: the C-ABI library loads, exports what the header
declares, and its host-only arithmetic (integer thresholds, folded Philox schedule, argument
validation) agrees with the oracle.  No compute call is made without a GPU."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(ig):
    hdr = open(os.path.join(ROOT, "include", "ising_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ising_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = C.CDLL(ig.lib_path())
    for name in sorted(declared):
        assert hasattr(lib, name), "libising_b200.so does not export " + name
    from ising_gpu_b200 import engine
    assert declared == set(engine.EXPORTED_SYMBOLS)
    assert lib.ising_abi_version() == 2


@pytest.mark.parametrize("temp", [1.5, 2.269185, 0.5, 0.3, 0.05 * 2.26918531421, 10.0, 1e9, 0.0, -1.0])
def test_integer_thresholds_equal_float_rule(ig, oracle, temp):
    """accept_count(p) is exactly the number of x with uniform(x) <= p: check both sides of the edge."""
    tab_gpu = ig.accept_table(temp)
    tab_cpu = oracle.exp_table(temp)[1]
    assert np.array_equal(tab_gpu, tab_cpu)
    for p in tab_cpu:
        n = ig.accept_count(p)
        assert 0 <= n <= 1 << 32
        if n > 0:
            assert oracle.uniform(n - 1) <= p
        if n < 1 << 32:
            assert not (oracle.uniform(n) <= p)


def test_init_threshold(ig, oracle):
    n = ig.accept_count(0.5, strict=True)
    assert oracle.uniform(n - 1) < np.float32(0.5) and not (oracle.uniform(n) < np.float32(0.5))
    assert n == 2147483584   # uniform < 0.5f  <=>  x < 2^31 - 64 (rounding of (float)x)


def test_survey_threshold_vectors(ig):
    t = ig.accept_table(1.5)
    assert ig.accept_count(t[3]) - 1 == 0x11c9aad0 and ig.accept_count(t[4]) - 1 == 0x013c678d
    t = ig.accept_table(2.269185)
    assert ig.accept_count(t[3]) - 1 == 0x2bec32a0 and ig.accept_count(t[4]) - 1 == 0x078932eb
    assert all(ig.accept_count(p) == 1 << 32 for p in t[:3])


def test_folded_philox_schedule_matches_plain_philox(ig, oracle):
    rng = random.Random(7)
    for _ in range(3000):
        X = rng.choice([2048, 4096, 65536, 1 << 20])
        it, color = rng.choice([0, 1, 2, 128, 10 ** 6, (1 << 27) + 5, (1 << 31) - 1]), rng.randrange(2)
        grow, word, q = rng.randrange(0, 1 << 20), rng.randrange(0, X // 64), rng.randrange(8)
        seed = rng.choice([ig.DEFAULT_SEED, rng.getrandbits(64), 1, (1 << 64) - 1])
        tid = (((grow // 16) * (X // 2048) + word // 32) * 256 + (grow % 16) * 16 + word % 16) & 0xffffffff
        ctr = 16 * (2 * it + color) + 8 * ((word % 32) // 16) + q
        exp = oracle.philox((ctr & 0xffffffff, ctr >> 32, tid, 0), (seed & 0xffffffff, seed >> 32))
        assert ig.site_random(seed, X, it, color, grow, word, q) == exp


def test_resolve_temperature(ig):
    assert ig.resolve_temperature(1.5, 0.3) == np.float32(1.5)                  # -t wins (main.cu:1465)
    assert ig.resolve_temperature(None, 1.0) == np.float32(2.26918531421)
    assert ig.resolve_temperature() == np.float32(0.1) * np.float32(2.26918531421)


def test_compute_calls_fail_loudly_without_gpu(ig):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ig.IsingError):
        ig.IsingEngine(2048, 16)


def test_numa_binding_is_a_no_op_without_a_gpu():
    """slab.bind_process_near_gpu must never raise: without a GPU / sysfs entry it changes nothing."""
    import os
    import torch
    from ising_gpu_b200.slab import bind_process_near_gpu
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    before = os.sched_getaffinity(0)
    assert bind_process_near_gpu(0) is None
    assert os.sched_getaffinity(0) == before


def test_shipped_kernels_are_blackwell_native():
    """cuobjdump of the built library: the update kernels stage tiles with TMA (UTMALDG) behind mbarriers (SYNCS), the
    tensor-core variant issues tcgen05.mma kind::i8 (UTCIMMA) and reads TMEM with tcgen05.ld (LDTM).  The same
    histograms are committed under profiles/sass_opcodes_*.txt by __graft_entry__.build()."""
    import shutil
    import sys
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_histogram
    hs = sass_histogram.histograms()
    for tag, need in sass_histogram.MUST.items():
        assert tag in hs, tag
        for op in need:
            assert any(k.startswith(op) for k in hs[tag]["ops"]), (tag, op)
    assert hs["tc_update_pipe_kernel"]["ops"]["UTCIMMA"] >= 9          # 4 + 4 k-steps + the edge-row k-step
    assert hs["update_color_kernel_fast"]["ops"]["UTMALDG"] == 5
    # instruction budget of the product update kernel: per thread = 128 sites = 32 Philox calls of
    # 16 multiplies (+ 2 x 2 for the two thread ids' constants, + addressing); a regression here is a slowdown
    k2 = hs["update_color_kernel2_fast"]["ops"]
    assert k2["UTMALDG"] == 5
    assert 512 <= k2["IMAD.WIDE.U32"] <= 520
    assert k2["LOP3"] <= 640 and k2["ISETP"] <= 290
    # no mma.sync / wgmma fallbacks anywhere in the variant kernel
    assert not any(k.startswith(("HMMA", "IMMA")) for k in hs["tc_update_pipe_kernel"]["ops"])


def test_nvidia_smi_topo_parser():
    """The fallback of the NUMA helper when sysfs reports no node for the GPU."""
    from ising_gpu_b200.slab import parse_nvidia_smi_topo
    txt = ("\tGPU0\tGPU1\tNIC0\tCPU Affinity\tNUMA Affinity\tGPU NUMA ID\n"
           "GPU0\t X \tNV18\tSYS\t0-3,64-67\t0\t\tN/A\n"
           "GPU1\tNV18\t X \tSYS\t32-35\t1\t\tN/A\n"
           "NIC0\tSYS\tSYS\t X \t\t\t\t\n\nLegend:\n  X    = Self\n")
    assert parse_nvidia_smi_topo(txt, 0) == ({0, 1, 2, 3, 64, 65, 66, 67}, 0)
    assert parse_nvidia_smi_topo(txt, 1) == ({32, 33, 34, 35}, 1)
    assert parse_nvidia_smi_topo(txt, 5) == (set(), None)
    assert parse_nvidia_smi_topo("garbage", 0) == (set(), None)

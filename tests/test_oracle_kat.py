"""CPU tests of the oracle against the reference's known-answer vectors (no GPU).

Vectors: Random123 Philox4x32-10 KATs; optimized/README.md:128-137 / :240 / :307 spin counts;
SURVEY appendix A (derived from the same restatement and, for the README sizes, equal to the
README).  The minutes-long vectors (README it 16/32, KAT-4) are pinned by oracle/pin_kat.py,
whose log is committed as oracle/pin_kat.log.
"""
import os

import numpy as np
import pytest

PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    # default seed 463463564571 -> key (e894c51b, 6b): tid 0 step 0 call 0; tid 5 step 2 call 0
    ((0, 0, 0, 0), (0xe894c51b, 0x0000006b), (0x999540fc, 0xe7871c3d, 0x7c93c9e1, 0xd78f45a6)),
    ((32, 0, 5, 0), (0xe894c51b, 0x0000006b), (0xbcf22290, 0xb7c1161a, 0xb52f8fec, 0x7f9669b5)),
]


@pytest.mark.parametrize("ctr,key,out", PHILOX_KAT)
def test_philox_known_answers(oracle, ctr, key, out):
    assert oracle.philox(ctr, key) == out


def test_uniform_endpoints_and_monotone(oracle):
    assert float(oracle.uniform(0)) == 2.0 ** -33
    assert float(oracle.uniform(0xffffffff)) == 1.0
    xs = np.concatenate([np.arange(0, 1 << 32, 9973 * 4099, dtype=np.uint64),
                         np.arange((1 << 32) - 4096, 1 << 32, dtype=np.uint64)])
    us = np.array([oracle.uniform(int(x)) for x in xs[:20000]], dtype=np.float32)
    assert np.all(np.diff(us) >= 0)


def test_exp_table_values(oracle):
    t = oracle.exp_table(1.5)
    assert float(t[1][3]).hex() == "0x1.1c9aac0000000p-4" and float(t[1][4]).hex() == "0x1.3c678c0000000p-8"
    assert np.array_equal(t[0], t[1][::-1])      # exp_h[0][n] == exp_h[1][4-n] bit for bit
    t = oracle.exp_table(2.269185)
    assert float(t[1][3]).hex() == "0x1.5f61940000000p-3" and float(t[1][4]).hex() == "0x1.e24cba0000000p-6"
    t0 = oracle.exp_table(0.0)
    assert list(t0[1]) == [8.0, 4.0, 0.5, -4.0, -8.0]


SMALL = {  # (X, Y, T) -> {it: (up, down)}   
    (2048, 16, 1.5): {0: (16475, 16293), 1: (16699, 16069), 2: (16776, 15992), 3: (16947, 15821), 4: (17161, 15607)},
    (2048, 16, 2.269185): {0: (16475, 16293), 1: (16586, 16182), 2: (16534, 16234), 3: (16651, 16117), 4: (16816, 15952)},
    (4096, 4096, 1.5): {0: (8392244, 8384972), 4: (8393075, 8384141), 8: (8398447, 8378769), 12: (8394525, 8382691),
                        16: (8390573, 8386643)},
    (4096, 4096, 2.269185): {0: (8392244, 8384972), 4: (8389218, 8387998), 8: (8389386, 8387830),
                             12: (8394197, 8383019), 16: (8400407, 8376809)},
}


@pytest.mark.parametrize("key", list(SMALL))
def test_small_lattice_counts(oracle, key):
    X, Y, T = key
    L = oracle.OracleLattice(X, Y)
    L.init()
    for it in range(0, max(SMALL[key]) + 1):
        if it:
            L.sweeps(1, T)
        if it in SMALL[key]:
            assert L.count() == SMALL[key][it], "it %d" % it


def test_readme_kat1_initial_counts(oracle):
    """optimized/README.md:128 -- 65536x65536, default seed: up 2147484090, down 2147483206."""
    assert oracle.init_count(65536, 65536) == (2147484090, 2147483206)


def test_slab_invariance(oracle):
    """N devices x Y rows == 1 device x N*Y rows (tid uses the global block row, main.cu:514)."""
    a = oracle.OracleLattice(2048, 64)
    a.init(); a.sweeps(2, 1.5)
    # the same lattice described as 4 slabs of 16 rows is the same object for the oracle (rows = ndev*Y);
    # what must hold is that row r's randoms do not depend on how rows are grouped: check through
    # the reference thread-id formula evaluated per slab
    X, Y, ndev = 2048, 16, 4
    gridx, gridy = X // 2048, Y // 16
    for dev in range(ndev):
        for by in range(gridy):
            for ty in (0, 7, 15):
                r = dev * Y + by * 16 + ty
                tid_ref = ((dev * gridy + by) * gridx + 0) * 256 + ty * 16 + 3
                tid_glob = ((r // 16) * gridx + 0) * 256 + (r % 16) * 16 + 3
                assert tid_ref == tid_glob


def test_dump_format(oracle, tmp_path):
    L = oracle.OracleLattice(2048, 16)
    L.init()
    f = tmp_path / "d.txt"
    L.dump(str(f))
    lines = f.read_text().split("\n")
    assert len(lines) == 17 and lines[-1] == "" and all(len(l) == 2048 for l in lines[:-1])
    lat = L.lattice()
    assert lines[0] == "".join(map(str, lat[0])) and lines[1] == "".join(map(str, lat[1]))
    b, w = L.plane(0), L.plane(1)
    assert lines[0][0] == str(b[0, 0]) and lines[0][1] == str(w[0, 0])     # even row: black first
    assert lines[1][0] == str(w[1, 0]) and lines[1][1] == str(b[1, 0])     # odd row: white first


def test_ref_layout_and_bits_roundtrip(oracle):
    L = oracle.OracleLattice(2048, 16)
    L.init()
    r0, bits0 = L.export_ref(0), L.export_bits(0)
    assert r0.shape == (16, 64) and bits0.shape == (16, 32)
    assert int(r0[0, 0] & 0xF) == int(L.plane(0)[0, 0]) and int((r0[0, 0] >> 4) & 0xF) == int(L.plane(0)[0, 1])
    assert np.all((r0 & np.uint64(0xEEEEEEEEEEEEEEEE)) == 0)
    M = oracle.OracleLattice(2048, 16)
    M.import_ref(0, r0); M.import_bits(1, L.export_bits(1))
    assert np.array_equal(M.plane(0), L.plane(0)) and np.array_equal(M.plane(1), L.plane(1))


def test_pin_log_is_green():
    log = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "pin_kat.log")
    txt = open(log).read()
    assert "FAIL" not in txt and "KAT-1 65536x65536 T=1.5 it  16" in txt and "KAT-4" in txt

"""CPU: the oracle against golden outputs of the REAL reference binary run on a B200
(tests/golden/reference_b200.*).  This is what pins the oracle beyond the README vectors:
T_c, other seeds, ragged shapes, sub-lattices, a temperature ramp -- spin counts at every
print point and the dumped lattice text, byte for byte."""
import numpy as np
import pytest

import golden_util as G

CPU_CASES = ["small_T1.5", "small_Tc", "rag_Tc", "sub_T1.5", "alpha_default", "seed_T2", "ramp", "mid_T1.5", "mid_Tc"]
MIN_TEMP = np.float32(0.05) * np.float32(2.26918531421)


@pytest.mark.parametrize("case", CPU_CASES)
def test_oracle_reproduces_reference_run(oracle, case, tmp_path):
    c = G.cases()[case]
    o = G.parse_args(c["args"])
    temp = np.float32(o["t"]) if o["t"] is not None else (np.float32(o["a"] if o["a"] is not None else 0.1) *
                                                         np.float32(2.26918531421))
    L = oracle.OracleLattice(o["x"], o["y"], seed=o["s"], xsl=o["xsl"], ysl=o["ysl"])
    L.init()
    got = [("Initial", None) + L.count()]
    files = G.extract_files(case, str(tmp_path / "gold")) if o["o"] else {}
    checked = 0

    def dump_and_compare(it, t):
        nonlocal checked
        name = "lattice_%dx%d_T_%f_IT_%08d_0.txt" % (o["y"], o["x"], float(t), it)
        assert name in files, name
        p = tmp_path / "mine.txt"
        L.dump(str(p))
        assert p.read_bytes() == open(files[name], "rb").read(), name
        checked += 1

    for j in range(o["n"]):
        L.sweeps(1, temp)
        if o["p"] and (j + 1) % o["p"] == 0:
            got.append(("step", j + 1) + L.count())
            if o["o"]:
                dump_and_compare(j + 1, temp)
        if o["u"] and (j + 1) % o["u"][1] == 0:                 # main.cu:1848-1849
            temp = max(MIN_TEMP, np.float32(temp + np.float32(o["u"][0])))
    got.append(("Final", o["n"]) + L.count())
    if o["o"]:
        dump_and_compare(o["n"], temp)
        assert checked >= len(files)
    assert got == G.parse_counts(c["lines"])


def test_golden_index_covers_the_configs():
    idx = G.cases()
    assert {"big_Tc", "big_T1.5", "big_sub_T1.5", "bonds_J0.25", "corr_Tc"} <= set(idx)
    # BASELINE.json config 3 (65536^2 at T_c) final counts of the reference on B200
    assert G.parse_counts(idx["big_Tc"]["lines"])[-1][1] == 128
    # and the reference on B200 reproduces its own README (optimized/README.md:137)
    assert G.parse_counts(idx["big_T1.5"]["lines"])[-1] == ("Final", 128, 2147678887, 2147288409)


@pytest.mark.parametrize("case", ["bonds_J0.25", "bonds_J0.5_Tc", "bonds_sub"])
def test_oracle_bonds_reproduce_reference(oracle, case, tmp_path):
    """-J: hamiltInitB_k/hamiltInitW_k + the bond XORs in the update (main.cu:153-331, :575-618)."""
    if case not in G.cases():
        pytest.skip("golden case not generated")
    c = G.cases()[case]
    o = G.parse_args(c["args"])
    L = oracle.OracleLattice(o["x"], o["y"], seed=o["s"], xsl=o["xsl"], ysl=o["ysl"])
    L.init(); L.init_bonds(o["J"])
    files = G.extract_files(case, str(tmp_path / "gold"))
    got = [("Initial", None) + L.count()]
    for j in range(o["n"]):
        L.sweeps(1, o["t"])
        if o["p"] and (j + 1) % o["p"] == 0:
            got.append(("step", j + 1) + L.count())
            name = "lattice_%dx%d_T_%f_IT_%08d_0.txt" % (o["y"], o["x"], o["t"], j + 1)
            L.dump(str(tmp_path / "m.txt"))
            assert (tmp_path / "m.txt").read_bytes() == open(files[name], "rb").read(), name
    got.append(("Final", o["n"]) + L.count())
    assert got == G.parse_counts(c["lines"])


@pytest.mark.parametrize("case", ["corr_sub", "corr_4096", "corr_8192_T2"])
def test_oracle_correlation_reproduces_reference_file(oracle, case, tmp_path):
    """-c: the corr_ file of the reference, line by line (format main.cu:1129-1135)."""
    if case not in G.cases():
        pytest.skip("golden case not generated")
    c = G.cases()[case]
    o = G.parse_args(c["args"])
    L = oracle.OracleLattice(o["x"], o["y"], seed=o["s"], xsl=o["xsl"], ysl=o["ysl"])
    L.init()
    files = G.extract_files(case, str(tmp_path / "gold"))
    ref_lines = open([p for n, p in files.items() if n.startswith("corr_")][0]).read().split("\n")
    k = 0
    for j in range(o["n"]):
        L.sweeps(1, o["t"])
        if o["p"] and (j + 1) % o["p"] == 0:
            mine = ("%10d" % (j + 1)) + "".join(" % -12G" % v for v in L.correlation(128))
            assert mine == ref_lines[k], "measure at iteration %d" % (j + 1)
            k += 1
    assert k > 0

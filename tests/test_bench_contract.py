"""CPU checks of bench.py's contract: both arms build the same `config`, the README known-answer tables agree with
the oracle's pinned ones, and the reference arm's CPU fallback prints one well-formed JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_config_is_a_pure_function_of_the_command_line():
    import bench
    a = bench.workload_config(65536, 1, bench.DEFAULT_SEED, 16)
    b = bench.workload_config(65536, 1, bench.DEFAULT_SEED, 16)
    assert a == b and set(a) == {"workload", "l2_flush", "variants"}
    assert "65536x65536" in a["workload"] and "T=1.5" in a["workload"]
    assert bench.workload_config(32768, 8, bench.DEFAULT_SEED, 16)["workload"].startswith(
        "checkerboard Metropolis, 262144x65536 lattice (32768 rows per GPU)")
    assert bench.default_rows(8) == 32768 and bench.default_rows(4) == 65536


def test_readme_tables_agree_with_the_pinned_oracle_vectors():
    import bench
    from oracle import pin_kat
    src, tab = bench.README_KAT[(1, 65536)]
    assert "README.md:128" in src
    for it in (16, 32):
        assert tab[it] == pin_kat.KAT1[it][0]
    # every table is the up-spin count of a lattice with an even number of spins: up + down = rows * 65536
    for (n, rows), (_, t) in bench.README_KAT.items():
        assert sorted(t) == [16, 32, 48, 64, 80, 96, 112, 128]
        assert all(0 < v < n * rows * 65536 for v in t.values())


def test_reference_arm_cpu_fallback_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ref-cpu", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "spin-flip attempts/ns" and d["unit"] == "flips/ns"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and set(d["config"]) == {"workload", "l2_flush", "variants"}

"""GPU: the cuIsing command-line host (C++ over the C ABI) against the golden transcripts and
dump files of the reference binary: same options -> same magnetization lines, same files."""
import os
import subprocess

import pytest

import golden_util as G

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "ising_gpu_b200", "cuIsing")

CASES = ["small_T1.5", "small_Tc", "rag_Tc", "sub_T1.5", "alpha_default", "seed_T2", "ramp", "mid_T1.5", "mid_Tc",
         "big_Tc", "big_T1.5", "big_sub_T1.5", "expser", "magn_stop", "bonds_J0.25", "bonds_J0.5_Tc", "bonds_sub",
         "corr_sub", "corr_4096", "corr_8192_T2"]


@pytest.mark.parametrize("case", CASES)
def test_cli_matches_reference_transcript(case, tmp_path):
    if case not in G.cases():
        pytest.skip("golden case not generated")
    c = G.cases()[case]
    r = subprocess.run([CLI] + c["args"], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-1000:]
    mine = [ln for ln in r.stdout.splitlines()
            if "magnetization" in ln or ln.startswith("Changing temperature") or ln.startswith("exp[")]
    assert mine == c["lines"]
    gold = G.extract_files(case, str(tmp_path / "gold"))
    for name, path in gold.items():
        assert (tmp_path / name).exists(), "missing dump " + name
        assert (tmp_path / name).read_bytes() == open(path, "rb").read(), name
    assert "flips/ns" in r.stdout


def test_cli_bad_dims_prints_usage_and_exits_zero(tmp_path):
    r = subprocess.run([CLI, "-x", "100"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0 and "Please specify an X dim multiple of 2048" in r.stderr and "Usage:" in r.stdout


def test_cli_checkpoint_save_load_continues_the_run(tmp_path):
    """cuIsing -n 3 --save + cuIsing -n 4 --load == cuIsing -n 7 (final counts and dump); the reference has no
    checkpoint, so the golden is the uninterrupted run of the same program (itself pinned above)."""
    base = ["-x", "4096", "-y", "64", "-t", "2.269185"]

    def final(args, cwd):
        os.makedirs(cwd, exist_ok=True)
        r = subprocess.run([CLI] + base + args, cwd=cwd, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-1000:]
        return [ln.split("(iter")[0] for ln in r.stdout.splitlines() if ln.startswith("Final")][0]
    whole = final(["-n", "7", "-o"], tmp_path / "whole")
    ck = str(tmp_path / "state.ckpt")
    final(["-n", "3", "--save", ck], tmp_path / "part1")
    cont = final(["-n", "4", "--load", ck, "-o"], tmp_path / "part2")
    assert cont == whole
    # the continued run numbers its own iterations from 1: compare the lattice dumps by content
    d1 = [p for p in os.listdir(tmp_path / "whole") if p.startswith("lattice_")]
    d2 = [p for p in os.listdir(tmp_path / "part2") if p.startswith("lattice_")]
    assert len(d1) == 1 and len(d2) == 1
    assert open(tmp_path / "whole" / d1[0], "rb").read() == open(tmp_path / "part2" / d2[0], "rb").read()


def test_cli_checkpoint_continues_ramp_print_schedule_and_bonds(tmp_path):
    """-u (temperature ramp), -p (print cadence) and -J (random bonds) run on the iteration number of the whole run
    and on the Hamiltonian / temperature stored in the checkpoint: 4 sweeps + save, load + 4 sweeps prints exactly the
    magnetization lines 5..8 of the uninterrupted 8-sweep run; a checkpoint written with -J is refused without it."""
    base = ["-x", "2048", "-y", "64", "-t", "2.0", "-u", "0.25,3", "-p", "2", "-J", "0.2"]

    def run(args, cwd, ok=True):
        os.makedirs(cwd, exist_ok=True)
        r = subprocess.run([CLI] + args, cwd=cwd, capture_output=True, text=True, timeout=300)
        assert (r.returncode == 0) == ok, r.stderr[-1000:]
        return r
    whole = G.parse_counts(run(base + ["-n", "8"], tmp_path / "whole").stdout.splitlines())
    ck = str(tmp_path / "ramp.ckpt")
    run(base + ["-n", "4", "--save", ck], tmp_path / "p1")
    r2 = run(base + ["-n", "4", "--load", ck], tmp_path / "p2")
    assert "Restored temperature 2.250000" in r2.stdout          # one ramp step (after sweep 3) had happened
    cont = G.parse_counts(r2.stdout.splitlines())
    steps = lambda cs: [(it, up, dn) for kind, it, up, dn in cs if kind == "step"]   # noqa: E731
    assert steps(cont) == [s for s in steps(whole) if s[0] > 4]
    assert [c[2:] for c in cont if c[0] == "Final"] == [c[2:] for c in whole if c[0] == "Final"]
    # same lattice, other Hamiltonian: refused
    nob = [a for a in base if a not in ("-J", "0.2")]
    bad = run(nob + ["-n", "1", "--load", ck], tmp_path / "p3", ok=False)
    assert "random bonds" in bad.stderr


def test_cli_corr_warns_when_reference_output_would_differ(tmp_path):
    """-c with X not a multiple of 4096: the reference's chunk loop reads past the row (main.cu:905); the CLI says so."""
    r = subprocess.run([CLI, "-x", "2048", "-y", "64", "-t", "2.269185", "-n", "2", "-p", "2", "-c"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "not a multiple of 4096" in r.stderr
    r = subprocess.run([CLI, "-x", "4096", "-y", "64", "-t", "2.269185", "-n", "2", "-p", "2", "-c"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "warning" not in r.stderr

"""GPU: the ising_tensorcore command-line host against the golden runs of the reference's
tensorcore program (printed magnetisation and final.txt, byte for byte via its sha256)."""
import hashlib
import json
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "ising_gpu_b200", "ising_tensorcore")
IDX = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_tc_b200.json")))["cases"]


@pytest.mark.parametrize("case", ["t_small", "t_rect", "t_hot", "t_mid", "t_mid_cold", "t_big_T1.5"])
def test_cli_matches_reference(case, tmp_path):
    c = IDX[case]
    r = subprocess.run([CLI] + c["args"], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-1000:]
    assert re.search(r"average magnetism \(absolute\): (\S+)", r.stdout).group(1) == c["magnetism"]
    assert "updates per ns" in r.stdout and "REPORT:" in r.stdout
    if c["files"]:
        assert hashlib.sha256((tmp_path / "final.txt").read_bytes()).hexdigest() == c["final_txt_sha256"]


@pytest.mark.parametrize("case,ngpus", [("t_small", 2), ("t_mid", 2), ("t_mid_cold", 4), ("t_mid_cold", 8)])
def test_cli_multi_gpu_matches_reference(case, ngpus, tmp_path):
    """-g N must reproduce the single-GPU golden run: the reference keeps global thread ids
    (tensorcore/main.cu:268, :293 launch offsets), so the result does not depend on N."""
    import torch
    c = IDX[case]
    nby = int(c["args"][c["args"].index("-y") + 1])
    if torch.cuda.device_count() < ngpus or nby % ngpus:
        pytest.skip("needs %d GPUs and nby %% ngpus == 0" % ngpus)
    r = subprocess.run([CLI] + c["args"] + ["-g", str(ngpus)], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-1000:]
    assert "nGPUs: %d" % ngpus in r.stdout
    assert re.search(r"average magnetism \(absolute\): (\S+)", r.stdout).group(1) == c["magnetism"]
    if c["files"]:
        assert hashlib.sha256((tmp_path / "final.txt").read_bytes()).hexdigest() == c["final_txt_sha256"]

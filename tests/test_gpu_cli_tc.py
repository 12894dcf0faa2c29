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

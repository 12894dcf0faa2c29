#!/usr/bin/env python
"""Turn gpurun_out/golden/ (written on the GPU box by tools/gen_golden_ref.sh, i.e. by the
UNMODIFIED reference binary oracle/_ref/cuIsing_ref running on a B200) into the committed
fixtures tests/golden/reference_b200.json + reference_b200_files.tgz."""
import io
import json
import os
import re
import tarfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "golden")
DST = os.path.join(ROOT, "tests", "golden")
os.makedirs(DST, exist_ok=True)
index = {}
tar = tarfile.open(os.path.join(DST, "reference_b200_files.tgz"), "w:gz", compresslevel=9)
SRCS = [SRC, os.path.join(ROOT, "gpurun_out", "golden_extra")]
for case, d in sorted((c, os.path.join(sd, c)) for sd in SRCS if os.path.isdir(sd) for c in os.listdir(sd)):
    out = open(os.path.join(d, "stdout.txt")).read()
    entry = {"args": open(os.path.join(d, "args.txt")).read().split(), "lines": [], "files": []}
    for ln in out.splitlines():
        if "magnetization" in ln or ln.startswith("Changing temperature") or ln.startswith("exp["):
            entry["lines"].append(ln)
        m = re.match(r"Kernel execution time for (\d+) update steps", ln)
        if m:
            entry["steps_done"] = int(m.group(1))
    for f in sorted(os.listdir(d)):
        if f.startswith("lattice_") or f.startswith("corr_"):
            p = os.path.join(d, f)
            if os.path.getsize(p) > 2_000_000:
                continue
            tar.add(p, arcname=case + "/" + f)
            entry["files"].append(f)
    index[case] = entry
tar.close()
json.dump({"generator": "tools/gen_golden_ref.sh on a B200 via gpurun; reference = optimized/main.cu + utils.c "
                        "compiled unmodified with nvcc 12.9 -O3 --use_fast_math for sm_100a (oracle/Makefile)",
           "cases": index}, open(os.path.join(DST, "reference_b200.json"), "w"), indent=1)
print({k: (len(v["lines"]), len(v["files"])) for k, v in index.items()})
print(os.path.getsize(os.path.join(DST, "reference_b200_files.tgz")))

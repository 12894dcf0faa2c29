#!/usr/bin/env python
"""Per-kernel launch counts, total time and share of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import json
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
tot = collections.OrderedDict()
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    k = r[ki][:90]
    t = tot.setdefault(k, [0, 0.0])
    t[0] += 1
    t[1] += float(r[vi].replace(",", "")) * 1e-6
allms = sum(t[1] for t in tot.values())
out = {"command": sys.argv[2] if len(sys.argv) > 2 else None,
       "kernels": {k: {"launches": n, "total_ms": round(ms, 4), "mean_us": round(ms / n * 1e3, 2), "share": round(ms / allms, 5)}
                   for k, (n, ms) in tot.items()}}
print(json.dumps(out, indent=1))

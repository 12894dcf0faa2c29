#!/usr/bin/env python
"""Top stall instructions of an `ncu -i X.ncu-rep --page source --csv` export, with the dominant stall reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
si, wi, ei = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(float(r[wi] or 0) for r in body)
print("total samples %d" % tot)
agg = {hdr[i]: sum(float(r[i] or 0) for r in body) for i in st}
print("by reason:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.005 * tot))
for idx, r in sorted(enumerate(body), key=lambda ir: -float(ir[1][wi] or 0))[:n]:
    why = sorted(((float(r[i] or 0), hdr[i][6:]) for i in st), reverse=True)[:2]
    print("%5.2f%%  #%-4d exec %-9s %-70s %s" % (100 * float(r[wi]) / tot, idx, r[ei], r[si].strip()[:70],
                                               " ".join("%s:%d" % (w, v) for v, w in why if v)))

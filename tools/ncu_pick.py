#!/usr/bin/env python
"""Print selected metrics of an `ncu -i X.ncu-rep --page raw --csv` export (first kernel)."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]; vals = rows[2] if len(rows) > 2 else rows[1]
pat = sys.argv[2] if len(sys.argv) > 2 else (
    r"issue_stalled.*per_issue_active|pipe_.*pct_of_peak_sustained_active$|inst_executed_pipe_[a-z0-9_]+\.sum$|"
    r"dram__bytes_(read|write)\.sum$|sm__cycles_elapsed.avg$|lts__t_bytes.sum$|sm__inst_executed.sum$|gpu__time_duration.sum|"
    r"smsp__cycles_active.avg$|issue_active.avg.pct|launch__registers|launch__occupancy_limit|sm__warps_active.avg.pct")
for k, v in zip(hdr, vals):
    if re.search(pat, k):
        try:
            if float(v.replace(',', '')) == 0: continue
        except ValueError:
            pass
        print("%-100s %s" % (k, v))

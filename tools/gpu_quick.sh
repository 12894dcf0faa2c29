#!/bin/bash
# quick A/B: parity tests + timing + pipe counters of the update kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python tools/drive.py -n 32 --reps 2 2>&1 | tee gpurun_out/quick_timing.txt
timeout 120 python tools/drive.py -n 32 --reps 1 -t 2.269185 2>&1 | tee -a gpurun_out/quick_timing.txt
M=sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed.sum,smsp__cycles_active.avg,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__warps_eligible.avg.per_cycle_active
timeout 600 ncu --metrics $M --clock-control none -k regex:update_color -s 2 -c 1 --csv --log-file gpurun_out/update_pipes.csv python tools/drive.py -n 2 --warm 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/update_pipes.csv')) if len(r)>10]
hdr=rows[0]; mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value')
for r in rows[1:]: print("%-80s %s" % (r[mi], r[vi]))
PY

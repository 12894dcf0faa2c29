#!/bin/bash
# Round 2: persistent update kernel vs the one-CTA-per-tile kernel (ISING_B200_KERNEL=tile), same library.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/r2_pytest_gpu.log
M=sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed.sum,smsp__cycles_active.avg,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__warps_eligible.avg.per_cycle_active,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__warps_active.avg.per_cycle_active
rm -f gpurun_out/r2_ab_persistent.txt
for k in tile persistent; do
  echo "== ISING_B200_KERNEL=$k" | tee -a gpurun_out/r2_ab_persistent.txt
  ISING_B200_KERNEL=$k timeout 120 python tools/drive.py -n 32 --reps 3 2>&1 | tee -a gpurun_out/r2_ab_persistent.txt
  ISING_B200_KERNEL=$k timeout 120 python tools/drive.py -n 32 --reps 2 -t 2.269185 2>&1 | tee -a gpurun_out/r2_ab_persistent.txt
  ISING_B200_KERNEL=$k timeout 600 ncu --metrics $M --clock-control none -k regex:update_color -s 2 -c 1 --csv --log-file gpurun_out/r2_pipes_$k.csv python tools/drive.py -n 2 --warm 1 > /dev/null 2>&1
  python - gpurun_out/r2_pipes_$k.csv <<'PY' | tee -a gpurun_out/r2_ab_persistent.txt
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value')
for r in rows[1:]: print("   %-85s %s" % (r[mi], r[vi]))
PY
done

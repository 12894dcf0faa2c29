#!/bin/bash
# A/B of update-kernel builds (ab/lib_*.so): throughput + spin counts (bit-exactness check) for all,
# ncu pipe counters for those named in $NCU_LIBS.
mkdir -p gpurun_out; out=gpurun_out/ab2.txt; rm -f $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee -a $out
M=sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed.sum,smsp__cycles_active.avg,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__warps_eligible.avg.per_cycle_active,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__warps_active.avg.per_cycle_active,sm__cycles_elapsed.avg.per_second
for lib in ab/lib_*.so; do
  echo "== $lib" | tee -a $out
  ISING_B200_LIB=$PWD/$lib timeout 120 python tools/drive.py -n 16 --reps 2 2>&1 | tee -a $out
done
for name in $NCU_LIBS; do
  lib=ab/lib_$name.so
  echo "== ncu $lib" | tee -a $out
  ISING_B200_LIB=$PWD/$lib timeout 600 ncu --metrics $M --clock-control none -k regex:update_color -s 2 -c 1 --csv --log-file gpurun_out/pipes_$name.csv python tools/drive.py -n 2 --warm 1 > /dev/null 2>&1
  python - gpurun_out/pipes_$name.csv <<'PY' | tee -a $out
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value')
for r in rows[1:]: print("   %-85s %s" % (r[mi], r[vi]))
PY
done

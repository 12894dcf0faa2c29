#!/bin/bash
mkdir -p gpurun_out
M=sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_uniform.sum,sm__inst_executed_pipe_adu.sum,sm__inst_executed_pipe_cbu.sum,sm__inst_executed.sum,smsp__cycles_active.avg,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
timeout 200 tools/pipe_map > gpurun_out/pipe_map.txt 2>&1
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/pipe_map_ncu.csv tools/pipe_map > /dev/null 2>&1
for lib in tools/libising_v0.so ising_gpu_b200/libising_b200.so; do
  echo "== $lib" >> gpurun_out/variants.txt
  ISING_B200_LIB=$PWD/$lib timeout 120 python tools/drive.py -n 32 --reps 2 >> gpurun_out/variants.txt 2>&1
  ISING_B200_LIB=$PWD/$lib timeout 120 python tools/drive.py -n 32 --reps 2 -t 2.269185 >> gpurun_out/variants.txt 2>&1
done
cat gpurun_out/variants.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 ncu --metrics $M,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:update_color -s 2 -c 1 --csv --log-file gpurun_out/update_pipes.csv python tools/drive.py -n 2 --warm 1 > /dev/null 2>&1

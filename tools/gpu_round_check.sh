#!/bin/bash
# Full check on a fresh box: GPU test suite, both bench arms, ncu launch list of the bench command,
# ncu --set full of the update kernel and of the tensor-core variant kernel.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; cut -c1-300 gpurun_out/bench_ref_n1.json
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:update_color -s 4 -c 1 -f -o gpurun_out/prof_update python tools/drive.py -n 4 --warm 1 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_update_pipe -s 4 -c 1 -f -o gpurun_out/prof_tc python tools/drive_tc.py -n 4 -w 2 > gpurun_out/ncu_full_tc.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 300 python tools/drive_tc.py -n 200 -w 20 | tee gpurun_out/tc_perf.log
timeout 300 python tools/drive_tc.py -n 200 -w 20 -a 1.0 | tee -a gpurun_out/tc_perf.log

#!/bin/bash
# Round-end evidence: bench line, ncu launch list of the same command, ncu --set full of the
# update kernel and of the tensor-core variant kernel, raw/source csv exports.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; cut -c1-250 gpurun_out/bench_final_n1.json
timeout 900 python bench.py --impl reference --steps 8 --warmup 1 > gpurun_out/bench_final_ref_n1.json 2>/dev/null; cut -c1-250 gpurun_out/bench_final_ref_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:update_color -s 4 -c 1 -o gpurun_out/prof_update_final python tools/drive.py -n 4 --warm 1 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_update_mma -s 4 -c 1 -o gpurun_out/prof_tc_final python tools/drive_tc.py -n 4 -w 2 > gpurun_out/ncu_full_tc.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 300 python tools/drive_tc.py -n 200 -w 20

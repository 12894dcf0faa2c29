#!/bin/bash
# Round 2 final evidence on one GPU: GPU test suite, both bench arms with the driver's command lines, ncu launch list of
# the bench command, edge/interior overlap timeline.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/r2_final_pytest_gpu_1gpu.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_ref_n1.json 2> gpurun_out/r2_final_bench_ref_n1.err; cut -c1-400 gpurun_out/r2_final_bench_ref_n1.json; echo
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err ) 2>&1 | tail -3; cat gpurun_out/r2_final_bench_n1.json; tail -3 gpurun_out/r2_final_bench_n1.err
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_ncu_launch_list_bench.csv $CMD > gpurun_out/r2_bench_under_ncu.log 2>&1
python tools/ncu_launch_share.py gpurun_out/r2_ncu_launch_list_bench.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 $CMD" | tee gpurun_out/r2_ncu_launch_share_bench.json
timeout 300 python tools/edge_overlap_timeline.py --rows 32768 --steps 6 2>&1 | tee gpurun_out/r2_edge_overlap_timeline.txt

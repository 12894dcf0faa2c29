#!/bin/bash
# Round 2, last pass on one GPU with the final kernels: test suite, both bench arms, ncu --set full of the two hot kernels
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r2_final_pytest_gpu_1gpu.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_ref_n1.json 2> gpurun_out/r2_final_bench_ref_n1.err; cut -c1-200 gpurun_out/r2_final_bench_ref_n1.json; echo
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; cut -c1-200 gpurun_out/r2_final_bench_n1.json; tail -3 gpurun_out/r2_final_bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:update_color -s 4 -c 1 -f -o gpurun_out/r2_prof_update_final python tools/drive.py -n 4 --warm 1 > gpurun_out/r2_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_update_pipe -s 4 -c 1 -f -o gpurun_out/r2_prof_tc_final python tools/drive_tc.py -n 4 -w 2 > gpurun_out/r2_ncu_full_tc.log 2>&1
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_ncu_launch_list_bench.csv $CMD > gpurun_out/r2_bench_under_ncu.log 2>&1
python tools/ncu_launch_share.py gpurun_out/r2_ncu_launch_list_bench.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 $CMD" > gpurun_out/r2_ncu_launch_share_bench.json
ls -la gpurun_out/r2_prof_*final*.ncu-rep
python -c "import __graft_entry__ as g; g.smoke()"

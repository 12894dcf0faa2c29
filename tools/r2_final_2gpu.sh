#!/bin/bash
# Round 2 final, two GPUs: GPU test suite, N=2 bench line, ncu launch list of cuIsing -d 2 (fused push, edge launch on
# its own stream: two kernels per colour and device).
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 ) 2>&1 | tee gpurun_out/r2_final_pytest_gpu_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n2.json 2> gpurun_out/r2_final_bench_n2.err; cut -c1-200 gpurun_out/r2_final_bench_n2.json; tail -2 gpurun_out/r2_final_bench_n2.err
timeout 900 python bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_ref_n2.json 2> gpurun_out/r2_final_bench_ref_n2.err; cut -c1-200 gpurun_out/r2_final_bench_ref_n2.json
CMD="ising_gpu_b200/cuIsing -x 65536 -y 32768 -d 2 -n 4 -t 1.5"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_ncu_launch_list_cuIsing_d2.csv $CMD > gpurun_out/r2_cuIsing_d2_under_ncu.log 2>&1
python tools/ncu_launch_share.py gpurun_out/r2_ncu_launch_list_cuIsing_d2.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 200 $CMD" | tee gpurun_out/r2_ncu_launch_share_cuIsing_d2.json
(cd gpurun_out && timeout 300 ../ising_gpu_b200/cuIsing -x 65536 -y 65536 -n 128 -p 16 -d 2 -t 1.5 > r2_cli_d2_kat2_transcript.txt 2>&1; tail -12 r2_cli_d2_kat2_transcript.txt)

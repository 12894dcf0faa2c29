#!/bin/bash
# 2-GPU check: full GPU test suite (multi-GPU cases un-skipped) and the N=2 bench line
mkdir -p gpurun_out
( time timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/pytest_gpu_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json; tail -2 gpurun_out/bench_n2.err
(cd /tmp && timeout 120 $GRAFT_REPO_ROOT/ising_gpu_b200/ising_tensorcore -x 128 -y 128 -g 2 -w 20 -n 200 -a 0.661030) 2>&1 | tee gpurun_out/tc_cli_g2.log | grep -E "updates per ns|magnetism"

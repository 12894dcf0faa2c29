#!/bin/bash
# multi-GPU checks: slab tests (needs >= 2 GPUs) + bench at N = 1, 2 (+4, 8 when available)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 900 python -m pytest tests/test_gpu_slabs.py -m gpu -x -q 2>&1 | tail -5
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 8 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
    tail -2 gpurun_out/bench_n$n.err; cat gpurun_out/bench_n$n.json
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 8 --warmup 3 --no-overlap --no-e2e > gpurun_out/bench_n${n}_noov.json 2> gpurun_out/bench_n${n}_noov.err
    cat gpurun_out/bench_n${n}_noov.json
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $n --steps 8 --warmup 1 > gpurun_out/bench_ref_n$n.json 2> gpurun_out/bench_ref_n$n.err
    cat gpurun_out/bench_ref_n$n.json
  fi
done
# single-process -d N CLI
if [ $N -ge 2 ]; then (cd gpurun_out && timeout 300 ../ising_gpu_b200/cuIsing -x 65536 -y 65536 -n 128 -p 16 -d 2 -t 1.5 > cli_d2.txt 2>&1; tail -12 cli_d2.txt); fi

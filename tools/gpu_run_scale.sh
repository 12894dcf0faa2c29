#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
for n in 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 8 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
    tail -2 gpurun_out/bench_n$n.err; cut -c1-400 gpurun_out/bench_n$n.json
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --impl reference --gpus $n --steps 8 --warmup 1 > gpurun_out/bench_ref_n$n.json 2> gpurun_out/bench_ref_n$n.err
    cut -c1-300 gpurun_out/bench_ref_n$n.json; echo
  fi
done
if [ $N -ge 8 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 8 --steps 8 --warmup 3 --rows-per-gpu 65536 --no-e2e > gpurun_out/bench_n8_kat3size.json 2> gpurun_out/bench_n8_kat3size.err
  cut -c1-400 gpurun_out/bench_n8_kat3size.json
  (cd gpurun_out && timeout 300 ../ising_gpu_b200/cuIsing -x 65536 -y 65536 -n 128 -p 16 -d 8 -t 1.5 > cli_d8.txt 2>&1; tail -12 cli_d8.txt)
fi

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slabs.py tests/test_gpu_bonds_corr.py -m gpu -x -q --timeout 300 2>&1 | tail -8
for halo in fused nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 8 --warmup 3 --halo $halo 2> gpurun_out/bench_n2_$halo.err > gpurun_out/bench_n2_$halo.json; tail -2 gpurun_out/bench_n2_$halo.err; cut -c1-200 gpurun_out/bench_n2_$halo.json; echo
done
(cd gpurun_out && timeout 300 ../ising_gpu_b200/cuIsing -x 65536 -y 65536 -n 128 -p 16 -d 2 -t 1.5 > cli_d2.txt 2>&1; tail -4 cli_d2.txt)

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/drive.py -n 64 --reps 2
for halo in fused nccl fused nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 8 --warmup 3 --halo $halo --no-e2e 2> /dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$halo', d['value'], d['ms_per_step'])"
done
for m in fused copy; do (cd gpurun_out && ISING_B200_HALO=$m timeout 300 ../ising_gpu_b200/cuIsing -x 65536 -y 65536 -n 128 -d 2 -t 1.5 | tail -1); done
timeout 900 python -m pytest tests/test_gpu_slabs.py tests/test_gpu_parity.py -m gpu -x -q --timeout 300 2>&1 | tail -3

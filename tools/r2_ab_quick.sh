#!/bin/bash
# quick A/B: parity smoke + throughput of both update kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py -m gpu -x -q 2>&1 | tail -3
for k in tile persistent; do
  echo "== ISING_B200_KERNEL=$k"
  ISING_B200_KERNEL=$k timeout 120 python tools/drive.py -n 32 --reps 3 2>&1
done 2>&1 | tee gpurun_out/r2_ab_quick.txt

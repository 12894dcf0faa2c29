#!/bin/bash
# quick A/B: parity + throughput of the update kernels (ISING_B200_KERNEL = tile1 | default (two tiles per CTA) | persistent)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py tests/test_gpu_cli_golden.py -m gpu -x -q 2>&1 | tail -3
for k in tile1 default; do
  echo "== ISING_B200_KERNEL=$k"
  ISING_B200_KERNEL=$k timeout 120 python tools/drive.py -n 32 --reps 3 2>&1
  ISING_B200_KERNEL=$k timeout 120 python tools/drive.py -n 32 --reps 1 -t 2.269185 2>&1
done 2>&1 | tee gpurun_out/r2_ab_quick.txt

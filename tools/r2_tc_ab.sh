#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r2_tc_ab.log
for lib in "$@"; do
  echo "== $lib" | tee -a gpurun_out/r2_tc_ab.log
  ISING_B200_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q 2>&1 | tail -2 | tee -a gpurun_out/r2_tc_ab.log
  for i in 1 2; do ISING_B200_LIB=$PWD/$lib timeout 300 python tools/drive_tc.py -n 200 -w 20; done 2>&1 | tee -a gpurun_out/r2_tc_ab.log
done

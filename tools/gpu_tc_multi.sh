#!/bin/bash
# 2-GPU check of the tensor-core variant (-g N): parity tests + updates/ns at -g 1 and -g 2.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_cli_tc.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/tc_multi_tests.log
cat gpurun_out/tc_multi_tests.log
for g in 1 2; do
  (cd /tmp && timeout 600 $GRAFT_REPO_ROOT/ising_gpu_b200/ising_tensorcore -x 128 -y 128 -w 20 -n 100 -a 0.661030 -g $g) 2>&1 | tail -12 > gpurun_out/tc_g${g}.log
  cat gpurun_out/tc_g${g}.log
done
(cd /tmp && timeout 600 $GRAFT_REPO_ROOT/ising_gpu_b200/ising_tensorcore -x 128 -y 256 -w 20 -n 100 -a 0.661030 -g 2) 2>&1 | tail -12 > gpurun_out/tc_g2_weak.log
cat gpurun_out/tc_g2_weak.log

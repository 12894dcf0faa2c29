#!/bin/bash
# variant iteration: parity tests, updates/ns at 32768^2, one ncu --set full capture of the MMA kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_cli_tc.py tests/test_golden_tc.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/tc_iter_tests.log
cat gpurun_out/tc_iter_tests.log
for rep in 1 2; do
(cd /tmp && timeout 600 $GRAFT_REPO_ROOT/ising_gpu_b200/ising_tensorcore -x 128 -y 128 -w 20 -n 200 -a 0.661030) 2>&1 | grep -E "updates per ns|magnetism" | tee -a gpurun_out/tc_iter_perf.log
done
(cd /tmp && timeout 600 $GRAFT_REPO_ROOT/ising_gpu_b200/ising_tensorcore -x 128 -y 128 -w 20 -n 200 -a 1.0) 2>&1 | grep -E "updates per ns|magnetism" | tee -a gpurun_out/tc_iter_perf.log
if [ "$1" = "ncu" ]; then
(cd /tmp && timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_update_mma -s 4 -c 1 -o $GRAFT_REPO_ROOT/gpurun_out/tc_variant_iter -f $GRAFT_REPO_ROOT/ising_gpu_b200/ising_tensorcore -x 128 -y 128 -w 2 -n 2 -a 0.661030) > gpurun_out/tc_iter_ncu.log 2>&1
tail -3 gpurun_out/tc_iter_ncu.log
fi

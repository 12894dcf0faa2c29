#!/bin/bash
# tensor-core variant: parity tests + updates/ns, persistent pipeline (default) vs one CTA per tile.
# Short timeouts and stop at the first failure: a hung kernel must not eat the GPU budget.
mkdir -p gpurun_out; out=gpurun_out/tc_pipe.txt; rm -f $out
timeout 60 python tools/drive_tc.py -x 8 -y 8 -n 4 -w 1 2>&1 | tee -a $out; [ ${PIPESTATUS[0]} -eq 0 ] || { echo "small run failed/hung" | tee -a $out; exit 1; }
timeout 150 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_cli_tc.py tests/test_golden_tc.py -m gpu -x -q 2>&1 | tail -5 | tee -a $out
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "tests failed/hung" | tee -a $out; exit 1; }
for k in pipe tile; do
  echo "== ISING_TC_KERNEL=$k" | tee -a $out
  ISING_TC_KERNEL=$k timeout 60 python tools/drive_tc.py -n 200 -w 20 2>&1 | tee -a $out
  ISING_TC_KERNEL=$k timeout 60 python tools/drive_tc.py -n 200 -w 20 -a 1.0 2>&1 | tee -a $out
done
if [ "$1" = "ncu" ]; then
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:tc_update_pipe -s 4 -c 1 -f -o gpurun_out/prof_tc_pipe python tools/drive_tc.py -n 4 -w 2 > gpurun_out/ncu_tc_pipe.log 2>&1
  tail -2 gpurun_out/ncu_tc_pipe.log
fi

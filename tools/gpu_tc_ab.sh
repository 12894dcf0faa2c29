#!/bin/bash
# A/B of tensor-core variant builds ab/lib_tc*.so: quick parity, then updates/ns (T=1.5 and T_c)
mkdir -p gpurun_out; out=gpurun_out/tc_ab.txt; rm -f $out
for lib in ab/lib_tc*.so; do
  echo "== $lib" | tee -a $out
  export ISING_B200_LIB=$PWD/$lib
  timeout 60 python tools/drive_tc.py -x 8 -y 8 -n 4 -w 1 2>&1 | tee -a $out; [ ${PIPESTATUS[0]} -eq 0 ] || { echo "small run failed/hung" | tee -a $out; continue; }
  timeout 150 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q 2>&1 | tail -2 | tee -a $out
  [ ${PIPESTATUS[0]} -eq 0 ] || { echo "tests failed/hung" | tee -a $out; continue; }
  timeout 60 python tools/drive_tc.py -n 200 -w 20 2>&1 | tee -a $out
  timeout 60 python tools/drive_tc.py -n 200 -w 20 -a 1.0 2>&1 | tee -a $out
  if [ -n "$NCU" ]; then
    timeout 200 ncu --set full --clock-control none --import-source on -k regex:tc_update_pipe -s 4 -c 1 -f -o gpurun_out/prof_$(basename $lib .so) python tools/drive_tc.py -n 4 -w 2 > /dev/null 2>&1
  fi
done

#!/bin/bash
# throughput + spin counts of update-kernel builds: tools/r2_ab_libs.sh ab/lib_a.so ab/lib_b.so ...
mkdir -p gpurun_out
for lib in "$@"; do
  echo "== $lib"
  ISING_B200_LIB=$PWD/$lib timeout 120 python tools/drive.py -n 32 --reps 2 2>&1
done 2>&1 | tee gpurun_out/r2_ab_libs.txt

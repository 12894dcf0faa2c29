#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_cli_tc.py -m gpu -x -q 2>&1 | tail -4
for i in 1 2; do timeout 300 python tools/drive_tc.py -n 200 -w 20; done 2>&1 | tee gpurun_out/r2_tc_perf.log
timeout 300 python tools/drive_tc.py -n 200 -w 20 -a 1.0 | tee -a gpurun_out/r2_tc_perf.log

#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1
tail -15 gpurun_out/pytest_gpu.txt
timeout 120 python tools/drive.py -n 32 --reps 3 > gpurun_out/quick_timing.txt 2>&1; cat gpurun_out/quick_timing.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python tools/drive.py -n 4 --warm 1 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:update_color -s 2 -c 2 -o gpurun_out/prof_update python tools/drive.py -n 2 --warm 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out

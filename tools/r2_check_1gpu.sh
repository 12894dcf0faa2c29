#!/bin/bash
# Round 2, one GPU: full GPU test suite, both bench arms (driver's command lines).
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/r2_pytest_gpu_1gpu.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_ref_n1.json 2> gpurun_out/r2_bench_ref_n1.err; cut -c1-1500 gpurun_out/r2_bench_ref_n1.json; tail -3 gpurun_out/r2_bench_ref_n1.err
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ) 2>&1 | tail -3; cat gpurun_out/r2_bench_n1.json; tail -5 gpurun_out/r2_bench_n1.err

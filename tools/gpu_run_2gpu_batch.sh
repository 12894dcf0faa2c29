#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_slabs.py -m gpu -x -q -k "replica_batch or per_gpu" 2>&1 | tail -6 | tee gpurun_out/pytest_slab_batch.log
[ ${PIPESTATUS[0]} -eq 0 ] || exit 1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print(d['value'], d['e2e'])"; tail -3 gpurun_out/bench_n2.err

#!/bin/bash
# Round 2, eight GPUs of one box: the whole GPU test suite with nothing skipped for lack of devices, the N=8 bench line
# of both arms, and the README KAT-3 lattice (524288 x 65536) through bench.py's in-run check.
mkdir -p gpurun_out
nvidia-smi -L | wc -l
( time timeout 2400 python -m pytest tests -m gpu -q -rs 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/r2_pytest_gpu_8gpu.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d.get("roofline") or {}; e=d.get("e2e") or {}
    print(sys.argv[1], "value", d.get("value"), "kernel_ms", r.get("kernel_ms"), "e2e", e.get("value"), (e.get("with_lattice_download") or {}).get("value"), (e.get("dependent_chain") or {}).get("value"), "counts", (d.get("counts_at_sweeps") or {}).get("up"), "selfcheck", (d.get("selfcheck") or {}).get("ok"), "kat", d.get("kat"), "numa", e.get("numa_of_rank0"), "runs", d.get("runs"))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex); print(open(sys.argv[1].replace(".json",".err")).read()[-2500:])
PY
}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; show gpurun_out/r2_bench_n8.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_ref_n8.json 2> gpurun_out/r2_bench_ref_n8.err; show gpurun_out/r2_bench_ref_n8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 8 --warmup 3 --rows-per-gpu 65536 --no-e2e > gpurun_out/r2_bench_n8_kat3.json 2> gpurun_out/r2_bench_n8_kat3.err; show gpurun_out/r2_bench_n8_kat3.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; show gpurun_out/r2_bench_n4.json

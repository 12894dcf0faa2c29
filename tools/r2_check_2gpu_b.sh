#!/bin/bash
# Round 2, two GPUs, second pass: multi-GPU tests with the two-tile kernel + edge stream, A/B of the edge stream.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/r2_pytest_gpu_2gpu_b.log
run_bench() { name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 20 --warmup 5 $BARGS > gpurun_out/r2b_bench_n2_$name.json 2> gpurun_out/r2b_bench_n2_$name.err
  python - gpurun_out/r2b_bench_n2_$name.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]; e=d.get("e2e") or {}
    print(sys.argv[1], "value", d["value"], "kernel_ms", r["kernel_ms"], "launches", d["gpu_launches"], "e2e", e.get("value"), (e.get("with_lattice_download") or {}).get("value"), (e.get("dependent_chain") or {}).get("value"), "counts", d["counts_at_sweeps"]["up"], "selfcheck", (d.get("selfcheck") or {}).get("ok"), "kat", (d.get("kat") or {}).get("ok"), "numa", e.get("numa_of_rank0"))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
}
BARGS="--no-e2e" run_bench edgestream ISING_B200_EDGE_STREAM=1
BARGS="--no-e2e" run_bench samestream ISING_B200_EDGE_STREAM=0
BARGS="--no-e2e" run_bench edgestream2 ISING_B200_EDGE_STREAM=1
BARGS="--no-e2e --rows-per-gpu 32768" run_bench edgestream_32k ISING_B200_EDGE_STREAM=1
BARGS="--no-e2e --rows-per-gpu 32768" run_bench samestream_32k ISING_B200_EDGE_STREAM=0
BARGS="" run_bench full X=1
BARGS="--staging wc --no-tc" run_bench full_wc X=1

#!/bin/bash
# Round 2, two GPUs: the multi-GPU part of the test suite, bench at N=2 (edge-launch variants), sanitizer runs of
# the fused halo push.
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/r2_pytest_gpu_2gpu.log
run_bench() { # name, env..., args
  name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 20 --warmup 5 $BARGS > gpurun_out/r2_bench_n2_$name.json 2> gpurun_out/r2_bench_n2_$name.err
  python - gpurun_out/r2_bench_n2_$name.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print(sys.argv[1], "value", d["value"], "kernel_ms", r["kernel_ms"], "launches", d["gpu_launches"], "e2e", (d.get("e2e") or {}).get("value"), "counts", d["counts_at_sweeps"], "selfcheck", d.get("selfcheck"), "kat", d.get("kat"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
}
BARGS="--no-e2e" run_bench split ISING_B200_SPLIT_EDGES=1
BARGS="--no-e2e" run_bench nosplit ISING_B200_SPLIT_EDGES=0
BARGS="--no-e2e" run_bench persistent ISING_B200_KERNEL=persistent
BARGS="" run_bench full X=1
timeout 900 python bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_ref_n2.json 2> gpurun_out/r2_bench_ref_n2.err; cut -c1-900 gpurun_out/r2_bench_ref_n2.json
# sanitizers on the fused halo push (single process, -d 2) and on the 1-GPU update kernel
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 ising_gpu_b200/cuIsing -x 2048 -y 64 -d 2 -n 3 -t 2.0 -p 1 > gpurun_out/r2_sanitizer_${tool}_d2.log 2>&1
  echo "== $tool -d 2: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_${tool}_d2.log | tail -1)"
done
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 env ISING_B200_KERNEL=persistent ising_gpu_b200/cuIsing -x 4096 -y 64 -d 2 -n 3 -t 2.0 -p 1 > gpurun_out/r2_sanitizer_racecheck_d2_persistent.log 2>&1
echo "== racecheck persistent -d 2: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_racecheck_d2_persistent.log | tail -1)"
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 ising_gpu_b200/ising_tensorcore -x 2 -y 2 -g 2 -w 1 -n 2 > gpurun_out/r2_sanitizer_memcheck_tc_g2.log 2>&1
echo "== memcheck tensorcore -g 2: $(grep -E 'ERROR SUMMARY' gpurun_out/r2_sanitizer_memcheck_tc_g2.log | tail -1)"

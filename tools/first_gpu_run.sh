#!/bin/bash
# First GPU visit: microbenchmark, parity tests, quick timing, and the reference binary.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 tools/pipe_microbench > gpurun_out/pipe_microbench.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1
tail -5 gpurun_out/pytest_gpu.txt
timeout 300 python - > gpurun_out/quick_timing.txt 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
import ising_gpu_b200 as ig
for temp in (1.5, 2.269185):
    with ig.IsingEngine(65536, 65536, temp=temp) as e:
        e.init_lattice(); e.sweeps(8)
        for rep in range(3):
            ms = e.sweeps_timed(32)
            print("T=%g 65536^2 32 sweeps: %.3f ms -> %.1f flips/ns" % (temp, ms, 65536.0*65536*32/(ms*1e6)), e.count_spins(), flush=True)
PY
cat gpurun_out/quick_timing.txt
if [ -x oracle/_ref/cuIsing_ref ]; then
  (cd gpurun_out && timeout 300 ../oracle/_ref/cuIsing_ref -x 65536 -y 65536 -n 128 -p 16 -t 1.5 > ref_65536_T1.5.txt 2>&1; tail -4 ref_65536_T1.5.txt
   timeout 300 ../oracle/_ref/cuIsing_ref -x 65536 -y 65536 -n 128 -p 16 -t 2.269185 > ref_65536_Tc.txt 2>&1; tail -3 ref_65536_Tc.txt)
fi

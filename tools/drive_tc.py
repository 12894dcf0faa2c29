#!/usr/bin/env python
"""Timing driver of the tensor-core variant (reference defaults: -w 100 -n 1000)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ising_gpu_b200.tensorcore import TensorcoreEngine
ap = argparse.ArgumentParser()
ap.add_argument("-x", type=int, default=128); ap.add_argument("-y", type=int, default=128)
ap.add_argument("-w", type=int, default=10); ap.add_argument("-n", type=int, default=50)
ap.add_argument("-a", type=float, default=0.661030); ap.add_argument("--path", type=int, default=0)
a = ap.parse_args()
with TensorcoreEngine(a.x, a.y, alpha=a.a, nn_path=a.path) as T:
    T.init(); T.update(a.w)
    ms = T.update_timed(a.n)
    print("tensorcore variant path %d: %dx%d, %d iterations: %.3f ms -> %.1f updates/ns, sum %d" %
          (a.path, T.nx, T.ny, a.n, ms, float(T.nx) * T.ny * a.n / (ms * 1e6), T.sum()), flush=True)

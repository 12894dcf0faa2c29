#!/usr/bin/env python
"""Minimal driver: init + N sweeps through the C ABI (used under ncu and for quick timings)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ising_gpu_b200 as ig  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("-x", type=int, default=65536)
ap.add_argument("-y", type=int, default=65536)
ap.add_argument("-n", type=int, default=8)
ap.add_argument("-t", type=float, default=1.5)
ap.add_argument("--warm", type=int, default=2)
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
with ig.IsingEngine(a.x, a.y, temp=a.t) as e:
    e.init_lattice()
    e.sweeps(a.warm)
    for _ in range(a.reps):
        ms = e.sweeps_timed(a.n)
        print("%dx%d T=%g %d sweeps: %.3f ms  %.1f flips/ns  counts %s" %
              (a.y, a.x, a.t, a.n, ms, float(a.x) * a.y * a.n / (ms * 1e6), e.count_spins()), flush=True)

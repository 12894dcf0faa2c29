#!/usr/bin/env python
"""Timeline of the edge / interior launches of one slab (what ising_update_color does on a context with the fused halo
push): the two launches back to back on ONE stream against the edge launch on a second, high-priority stream.
CUDA events are recorded in-stream around every launch and read back as offsets from a common origin, so the overlap of
the 64-CTA edge kernel with the interior grid is visible without a tracing tool (nsys is not in the image; ncu
serialises kernels).  Uses the public partial-update entry point on a slab context whose ghost rows the caller owns
(same kernels as the fused path, minus the push stores).
    python tools/edge_overlap_timeline.py [--rows 32768] [--steps 6]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ising_gpu_b200 as ig  # noqa: E402
from ising_gpu_b200.engine import F_NO_PEER_COPY, PART_EDGES, PART_INTERIOR  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=32768)
ap.add_argument("--steps", type=int, default=6, help="colour steps shown")
a = ap.parse_args()
X = 65536
torch.cuda.set_device(0)
main = torch.cuda.Stream()
lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
edge = torch.cuda.Stream(priority=-1)


def run(two_streams):
    eng = ig.IsingEngine(X, a.rows, temp=1.5, slab_first=0, slab_count=2, flags=F_NO_PEER_COPY)
    eng.set_stream(0, main.cuda_stream)
    eng.init_lattice()
    torch.cuda.synchronize()
    ev = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731
    origin = ev()
    rec = []
    for warm in (True, False):
        with torch.cuda.stream(main):
            origin.record(main)
        for k in range(a.steps):
            color, it = k & 1, 1 + k // 2
            e0, e1, i0, i1 = ev(), ev(), ev(), ev()
            if two_streams:
                pre = torch.cuda.Event()
                pre.record(main)
                edge.wait_event(pre)
                e0.record(edge)
                eng.update_color_part(0, color, it, PART_EDGES, edge.cuda_stream)
                e1.record(edge)
                i0.record(main)
                eng.update_color_part(0, color, it, PART_INTERIOR, main.cuda_stream)
                i1.record(main)
                main.wait_event(e1)
            else:
                e0.record(main)
                eng.update_color_part(0, color, it, PART_EDGES, main.cuda_stream)
                e1.record(main)
                i0.record(main)
                eng.update_color_part(0, color, it, PART_INTERIOR, main.cuda_stream)
                i1.record(main)
            if not warm:
                rec.append((k, e0, e1, i0, i1))
        torch.cuda.synchronize()
    out = []
    for k, e0, e1, i0, i1 in rec:
        out.append((k, origin.elapsed_time(e0) * 1e3, origin.elapsed_time(e1) * 1e3, origin.elapsed_time(i0) * 1e3,
                    origin.elapsed_time(i1) * 1e3))
    eng.close()
    return out


for two in (False, True):
    print("== %s  (%d rows x %d columns per slab; microseconds from the start of the sequence)" %
          ("edge launch on a second high-priority stream" if two else "edge and interior on one stream", a.rows, X))
    print("  step   edge start    edge end   interior start  interior end   step length")
    tl = run(two)
    prev_end = 0.0
    for k, e0, e1, i0, i1 in tl:
        end = max(e1, i1)
        print("  %4d   %10.1f  %10.1f   %14.1f  %12.1f   %11.1f" % (k, e0, e1, i0, i1, end - prev_end))
        prev_end = end
    print("  mean step length: %.1f us" % ((max(tl[-1][2], tl[-1][4]) - max(tl[0][2], tl[0][4])) / (len(tl) - 1)))

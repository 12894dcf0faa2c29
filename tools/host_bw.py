#!/usr/bin/env python
"""Platform ceiling for bench.py's e2e leg: aggregate host<->device bandwidth of all ranks copying at once, with the
buffer sizes of the e2e step (two planes up, two planes down per step) and nothing else running.
    torchrun --nproc-per-node N tools/host_bw.py [--mib 128] [--iters 20] [--wc]
Prints one JSON line on rank 0: GB/s (sum over ranks) for H2D alone, D2H alone and both directions at once."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=128, help="MiB per plane (128 = 32768 rows x 65536 columns at 1 bit/spin)")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--wc", action="store_true", help="write-combined source buffers (ising_host_alloc)")
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = a.mib << 20
if a.wc:
    from ising_gpu_b200.engine import HostBuffer
    keep = [HostBuffer((n // 4,), write_combined=True) for _ in range(2)]
    src = [torch.from_numpy(h.array.view("int32")) for h in keep]
else:
    src = [torch.empty(n // 4, dtype=torch.int32).pin_memory() for _ in range(2)]
dst = [torch.empty(n // 4, dtype=torch.int32).pin_memory() for _ in range(2)]
for s in src:
    s.fill_(1)
dev_in = [torch.empty(n // 4, dtype=torch.int32, device="cuda") for _ in range(2)]
dev_out = [torch.ones(n // 4, dtype=torch.int32, device="cuda") for _ in range(2)]
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down):
    def once():
        if up:
            with torch.cuda.stream(s_up):
                for d, s in zip(dev_in, src):
                    d.copy_(s, non_blocking=True)
        if down:
            with torch.cuda.stream(s_dn):
                for d, s in zip(dst, dev_out):
                    d.copy_(s, non_blocking=True)
    once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.iters):
        once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return 2 * n * a.iters * world * (int(up) + int(down)) / float(t.item()) / 1e9


res = {"ranks": world, "mib_per_plane": a.mib, "write_combined_src": a.wc, "h2d_gbs": round(run(True, False), 1),
       "d2h_gbs": round(run(False, True), 1), "both_gbs": round(run(True, True), 1)}
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()

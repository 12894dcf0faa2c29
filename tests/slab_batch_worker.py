"""torchrun worker for test_gpu_slabs.py::test_slab_replica_batch: every rank runs a replica batch on its
slab (ising_run_slab_batch_from_host, two contexts per rank) and checks it against the plain
import -> refresh ghosts -> sweeps -> export path on the same inputs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ising_gpu_b200.slab import SlabEngine  # noqa: E402

X, Y, sweeps, nrep, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
temp, seed = 2.269185, 463463564571
eng = SlabEngine(X, Y, temp, seed, rank=rank, world=world, device=local, halo="fused")
with torch.cuda.stream(torch.cuda.Stream()):
    eng2 = SlabEngine(X, Y, temp, seed, rank=rank, world=world, device=local, halo="fused")
with torch.cuda.stream(torch.cuda.Stream()):
    eng3 = SlabEngine(X, Y, temp, seed, rank=rank, world=world, device=local, halo="fused")
eng.init_lattice()
eng2.init_lattice()
eng3.init_lattice()
rng = np.random.default_rng(100 + rank)
reps, want = [], []
for k in range(nrep):       # different lattices per replica: the plain path gives ghosts and expected results
    b = rng.integers(0, 2**32, size=(Y, X // 64), dtype=np.uint32)
    w = rng.integers(0, 2**32, size=(Y, X // 64), dtype=np.uint32)
    eng.eng.import_bits(b, w)
    eng.refresh_ghosts()
    g = eng.white_ghost_rows()
    eng.it = 0
    eng.sweeps(sweeps)
    torch.cuda.synchronize()
    want.append(eng.eng.export_bits() + (eng.count_spins(),))
    reps.append((b, w, g))
ok = True
for other in (eng2, [eng2, eng3]):          # two contexts per rank, then three (ising_run_batch_from_host_n)
    eng.it = eng2.it = eng3.it = 0
    torch.cuda.synchronize()
    dist.barrier()
    ob = [np.zeros((Y, X // 64), np.uint32) for _ in reps]
    ow = [np.zeros((Y, X // 64), np.uint32) for _ in reps]
    counts = eng.run_batch_from_host(other, [r[0] for r in reps], [r[1] for r in reps], [r[2] for r in reps], sweeps, ob, ow)
    ok = ok and all(np.array_equal(ob[k], want[k][0]) and np.array_equal(ow[k], want[k][1]) and counts[k] == want[k][2]
                    for k in range(nrep))
open(os.path.join(out, "ok%d" % rank), "w").write("1" if ok else "0")
dist.barrier()
dist.destroy_process_group()

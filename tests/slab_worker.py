"""torchrun worker for tests/test_gpu_slabs.py: one rank = one GPU = one slab."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ising_gpu_b200.slab import SlabEngine  # noqa: E402

X, Y, sweeps, temp, overlap, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), \
    int(sys.argv[5]), sys.argv[6]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
halo = "fused" if overlap == 2 else "nccl"
eng = SlabEngine(X, Y, temp, 463463564571, rank=rank, world=world, device=local, overlap=bool(overlap), halo=halo)
eng.init_lattice()
eng.sweeps(sweeps)
torch.cuda.synchronize()
b, w = eng.eng.export_bits()
np.save(os.path.join(out, "b%d.npy" % rank), b)
np.save(os.path.join(out, "w%d.npy" % rank), w)
dist.barrier()
dist.destroy_process_group()

"""torchrun worker for tests/test_gpu_slabs.py::test_whole_lattice_observables_across_processes: one rank = one GPU =
one slab; the observables of the WHOLE lattice (two-point correlation, bond sum) asked from every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ising_gpu_b200.slab import SlabEngine  # noqa: E402

X, Y, sweeps, temp, overlap, ncorr, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), \
    int(sys.argv[5]), int(sys.argv[6]), sys.argv[7]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
halo = "fused" if overlap == 2 else "nccl"
eng = SlabEngine(X, Y, temp, 463463564571, rank=rank, world=world, device=local, overlap=bool(overlap), halo=halo)
eng.init_lattice()
eng.sweeps(sweeps)
bs = eng.bond_sum()
corr = eng.correlation(ncorr, root=world - 1)       # a root other than rank 0: the result still reaches every rank
e = eng.energy_per_site()
eng.sweeps(1)                                       # the observables must leave the slab engine usable
torch.cuda.synchronize()
b, w = eng.eng.export_bits()
np.savez(os.path.join(out, "r%d.npz" % rank), b=b, w=w, corr=corr, bond_sum=np.int64(bs), energy=np.float64(e))
dist.barrier()
eng.close()
dist.destroy_process_group()

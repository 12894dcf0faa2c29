"""world_size-2 (and 3) gloo tests of the multi-process halo protocol on CPU.

The transport/ordering code under test is ising_gpu_b200.slab.HaloExchange -- the same class the
GPU path uses over NCCL.  The per-slab compute is the oracle in slab mode (test infrastructure),
so the check is: N slabs + HaloExchange == one full lattice, bit for bit (the N-GPU == 1-GPU
property of the reference, optimized/main.cu:514-515 + README runs with -d 2 / -d 8)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleSlab:
    def __init__(self, oracle, X, Y, rank):
        self.L = oracle.OracleLattice(X, Y)
        self.L.make_slab(rank * Y)
        self.Y = Y

    def edge_rows(self, color):
        p = self.L.plane(color)
        return torch.from_numpy(p[0]), torch.from_numpy(p[self.Y - 1])

    def ghost_rows(self, color):
        return torch.from_numpy(self.L.ghost(color, 0)), torch.from_numpy(self.L.ghost(color, 1))


def _worker(rank, world, port, X, Y, sweeps, temp, q):
    sys.path.insert(0, ROOT)
    from oracle import ising_oracle as oracle
    from ising_gpu_b200.slab import HaloExchange
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        slab = OracleSlab(oracle, X, Y, rank)
        halo = HaloExchange(rank, world)
        slab.L.init()
        for color in (0, 1):
            HaloExchange.finish(halo.start(slab, color))
        for it in range(1, sweeps + 1):
            for color in (0, 1):
                slab.L.update_color(color, it, temp)
                HaloExchange.finish(halo.start(slab, color))
        q.put((rank, slab.L.plane(0).copy(), slab.L.plane(1).copy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_slabs_plus_halo_equal_full_lattice(oracle, world):
    X, Y, sweeps, temp = 2048, 32, 3, 2.269185
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, X, Y, sweeps, temp, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        r, b, w = q.get(timeout=120)
        got[r] = (b, w)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = oracle.OracleLattice(X, Y * world)
    full.init()
    full.sweeps(sweeps, temp)
    for color in (0, 1):
        stacked = np.vstack([got[r][color] for r in range(world)])
        assert np.array_equal(stacked, full.plane(color))


def test_halo_is_noop_for_single_slab():
    from ising_gpu_b200.slab import HaloExchange
    assert HaloExchange(0, 1).start(None, 0) == []


def _gather_worker(rank, world, port, root, q):
    sys.path.insert(0, ROOT)
    from ising_gpu_b200.slab import gather_slabs
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        b = rng.integers(0, 2**32, size=(16, 32), dtype=np.uint32)
        w = rng.integers(0, 2**32, size=(16, 32), dtype=np.uint32)
        B, W = gather_slabs(b, w, rank, world, root=root)
        q.put((rank, b, w, B, W))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,root", [(2, 0), (3, 1)])
def test_gather_slabs_stacks_the_ranks_in_slab_order(world, root):
    """What SlabEngine.correlation feeds to the whole-lattice context: rank r's planes are rows [r*Y, (r+1)*Y) of the
    arrays the root gets (optimized/main.cu:1542-1545), nobody else gets anything."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, root, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        r, b, w, B, W = q.get(timeout=120)
        got[r] = (b, w, B, W)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        if r != root:
            assert got[r][2] is None and got[r][3] is None
    assert got[root][2].dtype == np.uint32
    assert np.array_equal(got[root][2], np.vstack([got[r][0] for r in range(world)]))
    assert np.array_equal(got[root][3], np.vstack([got[r][1] for r in range(world)]))


def test_gather_slabs_single_rank_is_identity():
    from ising_gpu_b200.slab import gather_slabs
    b = np.arange(64, dtype=np.uint32).reshape(2, 32)
    B, W = gather_slabs(b, b + 1, 0, 1)
    assert np.array_equal(B, b) and np.array_equal(W, b + 1)

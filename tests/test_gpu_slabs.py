"""GPU tests of the slab decomposition: ghost rows, edge/interior split, single-process -d N and
one-process-per-GPU over NCCL.  Property under test (optimized/main.cu:514-515, README -d runs):
N slabs of Y rows give exactly the lattice of one device with N*Y rows."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TC = 2.269185


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _oracle_planes(oracle, X, rows, sweeps, temp):
    L = oracle.OracleLattice(X, rows)
    L.init()
    L.sweeps(sweeps, temp)
    return L.export_bits(0), L.export_bits(1), L.count()


@pytest.mark.parametrize("nslab,Y", [(2, 32), (3, 16), (4, 48)])
def test_slabs_on_one_gpu_with_manual_halo(ig, oracle, nslab, Y):
    """Several slab contexts on device 0, ghost rows copied by the test through ising_row_ptr:
    exercises the ghost-row addressing and the EDGES / INTERIOR launches without needing 2 GPUs."""
    import torch
    from ising_gpu_b200.engine import F_NO_PEER_COPY, PART_EDGES, PART_INTERIOR
    from ising_gpu_b200.slab import _DevRows
    X, sweeps = 4096, 3
    engs = [ig.IsingEngine(X, Y, temp=TC, slab_first=s, slab_count=nslab, flags=F_NO_PEER_COPY) for s in range(nslab)]
    dev = torch.device("cuda", 0)

    def row(e, color, r):
        p, n = e.row_ptr(0, color, r)
        return torch.as_tensor(_DevRows(p, n), device=dev)

    def exchange(color):
        for e in engs:
            e.sync()
        for s, e in enumerate(engs):
            up, dn = engs[(s - 1) % nslab], engs[(s + 1) % nslab]
            row(up, color, Y).copy_(row(e, color, 0))          # my first row -> bottom ghost of the slab above
            row(dn, color, -1).copy_(row(e, color, Y - 1))     # my last row  -> top ghost of the slab below
        torch.cuda.synchronize()

    for e in engs:
        e.init_lattice()
    exchange(0); exchange(1)
    for it in range(1, sweeps + 1):
        for color in (0, 1):
            for e in engs:
                e.update_color_part(0, color, it, PART_EDGES)
                e.update_color_part(0, color, it, PART_INTERIOR)
            exchange(color)
    ob, ow, cnt = _oracle_planes(oracle, X, Y * nslab, sweeps, TC)
    gb = np.vstack([e.export_bits()[0] for e in engs])
    gw = np.vstack([e.export_bits()[1] for e in engs])
    assert np.array_equal(gb, ob) and np.array_equal(gw, ow)
    ups = sum(e.count_spins()[0] for e in engs)
    assert ups == cnt[0]
    for e in engs:
        e.close()


@pytest.mark.parametrize("ndev,mode", [(2, "fused"), (2, "copy"), (4, "fused"), (8, "fused"), (8, "copy")])
def test_single_process_n_devices(ig, oracle, ndev, mode, monkeypatch):
    """cuIsing -d N path: one context, N GPUs; halo pushed by the kernel (default) or copied.  Y = 96 rows per device
    gives 6 tile rows: the edge launch on its own stream and the two-tiles-per-CTA interior both run."""
    if _ngpu() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    monkeypatch.setenv("ISING_B200_HALO", mode)
    X, Y, sweeps = 4096, 96, 4
    with ig.IsingEngine(X, Y, ndev=ndev, temp=TC) as eng:
        eng.init_lattice()
        eng.sweeps(sweeps)
        ob, ow, cnt = _oracle_planes(oracle, X, ndev * Y, sweeps, TC)
        gb = np.vstack([eng.export_bits(d)[0] for d in range(ndev)])
        gw = np.vstack([eng.export_bits(d)[1] for d in range(ndev)])
        assert np.array_equal(gb, ob) and np.array_equal(gw, ow)
        assert eng.count_spins() == cnt
        assert abs(eng.energy_per_site() - (-2.0 * eng.correlation(1)[0])) < 1e-12   # ghost rows current on every device


@pytest.mark.parametrize("nranks,overlap", [(2, 1), (2, 0), (2, 2), (4, 2), (8, 2), (8, 1)])
def test_one_process_per_gpu_nccl(oracle, tmp_path, nranks, overlap):
    """torchrun, N ranks; halo over NCCL (1 overlapped, 0 serialised) or pushed by the update kernel
    itself through CUDA-IPC peer memory (2): stacked slabs == oracle."""
    if _ngpu() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    X, Y, sweeps = 4096, 96, 5
    out = tmp_path / "slabs"
    out.mkdir()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % nranks, "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "slab_worker.py"), str(X), str(Y),
           str(sweeps), str(TC), str(overlap), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    ob, ow, cnt = _oracle_planes(oracle, X, nranks * Y, sweeps, TC)
    gb = np.vstack([np.load(out / ("b%d.npy" % k)) for k in range(nranks)])
    gw = np.vstack([np.load(out / ("w%d.npy" % k)) for k in range(nranks)])
    assert np.array_equal(gb, ob) and np.array_equal(gw, ow)


def test_more_than_65535_tile_rows(ig):
    """1 120 000 rows = 70 000 tile rows (beyond gridDim.y): one tall context == two stacked slabs."""
    import torch
    from ising_gpu_b200.engine import F_NO_PEER_COPY, PART_ALL
    from ising_gpu_b200.slab import _DevRows
    X, Y2 = 2048, 560000
    dev = torch.device("cuda", 0)
    with ig.IsingEngine(X, 2 * Y2, temp=TC) as tall:
        tall.init_lattice()
        tall.sweeps(2)
        tb, tw = tall.export_bits()
        cnt = tall.count_spins()
    engs = [ig.IsingEngine(X, Y2, temp=TC, slab_first=s, slab_count=2, flags=F_NO_PEER_COPY) for s in range(2)]

    def row(e, color, r):
        p, n = e.row_ptr(0, color, r)
        return torch.as_tensor(_DevRows(p, n), device=dev)

    def exchange(color):
        for e in engs:
            e.sync()
        for s, e in enumerate(engs):
            o = engs[1 - s]
            row(o, color, Y2).copy_(row(e, color, 0))
            row(o, color, -1).copy_(row(e, color, Y2 - 1))
        torch.cuda.synchronize()

    for e in engs:
        e.init_lattice()
    exchange(0); exchange(1)
    for it in (1, 2):
        for color in (0, 1):
            for e in engs:
                e.update_color_part(0, color, it, PART_ALL)
            exchange(color)
    gb = np.vstack([e.export_bits()[0] for e in engs])
    gw = np.vstack([e.export_bits()[1] for e in engs])
    assert np.array_equal(gb, tb) and np.array_equal(gw, tw)
    assert sum(e.count_spins()[0] for e in engs) == cnt[0]
    for e in engs:
        e.close()


def test_slab_replica_batch(tmp_path):
    """torchrun, 2 ranks x 2 contexts: ising_run_slab_batch_from_host (5 different replicas, host ghosts,
    fused halo between the ranks, copies overlapping kernels) == the plain import/refresh/sweeps/export path."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    out = tmp_path / "batch"
    out.mkdir()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "tests", "slab_batch_worker.py"), "4096", "128",
           "4", "5", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert [(out / ("ok%d" % k)).read_text() for k in range(2)] == ["1", "1"]


def test_slab_engine_observables_single_rank(ig):
    """SlabEngine with one rank: whole-lattice observables are the context's own."""
    import torch
    from ising_gpu_b200.slab import SlabEngine
    X, Y = 4096, 64
    se = SlabEngine(X, Y, TC, 463463564571, rank=0, world=1, device=0)
    se.init_lattice()
    se.sweeps(3)
    torch.cuda.synchronize()
    with ig.IsingEngine(X, Y, temp=TC) as one:
        one.init_lattice()
        one.sweeps(3)
        assert np.array_equal(se.correlation(8), one.correlation(8))
        assert se.bond_sum() == one.bond_sum() and se.energy_per_site() == one.energy_per_site()
    se.close()


@pytest.mark.parametrize("nranks,overlap", [(2, 1), (2, 2), (4, 2)])
def test_whole_lattice_observables_across_processes(ig, tmp_path, nranks, overlap):
    """torchrun, N ranks (NCCL halo / fused push): SlabEngine.correlation and .bond_sum of the lattice spread over the
    processes == ising_correlation / ising_bond_sum of one context holding the stacked lattice (the reference's -c
    covers all devices, optimized/main.cu:1072-1138), on every rank; the run goes on bit-exactly afterwards."""
    if _ngpu() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    X, Y, sweeps, ncorr = 4096, 96, 4, 12
    out = tmp_path / "obs"
    out.mkdir()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % nranks, "--master-addr",
           "127.0.0.1", "--master-port", "29535", os.path.join(ROOT, "tests", "slab_obs_worker.py"), str(X), str(Y),
           str(sweeps), str(TC), str(overlap), str(ncorr), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = [np.load(out / ("r%d.npz" % k)) for k in range(nranks)]
    with ig.IsingEngine(X, Y * nranks, temp=TC) as one:
        one.init_lattice()
        one.sweeps(sweeps)
        corr, bs, e = one.correlation(ncorr), one.bond_sum(), one.energy_per_site()
        one.sweeps(1)
        ob, ow = one.export_bits()
    for g in got:
        assert np.array_equal(g["corr"], corr)
        assert int(g["bond_sum"]) == bs and float(g["energy"]) == e
    assert np.array_equal(np.vstack([g["b"] for g in got]), ob) and np.array_equal(np.vstack([g["w"] for g in got]), ow)

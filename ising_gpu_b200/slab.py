"""One-process-per-GPU slab decomposition: the lattice is cut into `world` slabs of Y rows
(rows [r*Y, (r+1)*Y) on rank r, ring-periodic), exactly the reference's -d decomposition
(optimized/main.cu:1542-1545, :1773; ring wrap slY = Y*ndev :1459).  The reference lets edge
blocks read the neighbour GPU's boundary row through managed memory (loadTile, :413-428) behind
a host barrier per colour (:1779-1784); here every slab keeps two ghost rows per colour, and
after each colour update the freshly written first/last row travels to the neighbours' ghost
rows (one X/16-byte message per direction) while the interior tiles are still being updated.

`HaloExchange` is transport + ordering only and is backend-agnostic: a backend exposes
    edge_rows(color)  -> (first_row, last_row)    1-D tensors to send
    ghost_rows(color) -> (top_ghost, bottom_ghost) 1-D tensors to receive into
    update(color, it, part)                         part in {ALL, EDGES, INTERIOR}
On GPUs the tensors alias the library's device memory and travel over NCCL (NVLink); the CPU
tests drive the same class over gloo with the oracle as backend.
"""
import torch
import torch.distributed as dist

from .engine import F_NO_PEER_COPY, PART_ALL, PART_EDGES, PART_INTERIOR, IsingEngine


def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def parse_nvidia_smi_topo(text, gpu):
    """(cpus, numa_node or None) of physical GPU `gpu` from the matrix `nvidia-smi topo -m` prints: tab-separated,
    one row per GPU, the columns "CPU Affinity" and "NUMA Affinity" after the link matrix."""
    import re
    lines = [re.sub(r"\x1b\[[0-9;]*m", "", ln) for ln in text.splitlines() if ln.strip()]
    hdr = next((ln for ln in lines if "CPU Affinity" in ln), None)
    row = next((ln for ln in lines if re.match(r"GPU%d\s" % gpu, ln)), None)
    if hdr is None or row is None:
        return set(), None
    cols = [c.strip() for c in hdr.split("\t")]
    cells = [c.strip() for c in row.split("\t")]
    if cols and cols[0] != "":          # no (empty) header cell above the row labels: shift
        cols = [""] + cols
    def cell(name):
        return cells[cols.index(name)] if name in cols and cols.index(name) < len(cells) else ""
    try:
        cpus = _cpulist(cell("CPU Affinity"))
    except ValueError:
        cpus = set()
    n = cell("NUMA Affinity")
    node = int(re.split(r"[,-]", n)[0]) if n[:1].isdigit() else None
    return cpus, node


def gpu_numa_topology(device):
    """Where the GPU hangs in the host: {"nodes": NUMA nodes of the box, "gpu_node": node of the GPU or None,
    "cpus": CPUs next to it, "source": where that came from}.  sysfs first (numa_node of the PCI function); when
    the platform reports -1 there, the "CPU Affinity" / "NUMA Affinity" columns of `nvidia-smi topo -m`."""
    import glob
    import os
    import subprocess
    info = {"nodes": len(glob.glob("/sys/devices/system/node/node[0-9]*")), "gpu_node": None, "cpus": set(), "source": None}
    try:
        pr = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node >= 0:
            info.update(gpu_node=node, source="sysfs",
                        cpus=_cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read()))
            return info
    except Exception:
        pass
    try:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[device]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else device
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        cpus, node = parse_nvidia_smi_topo(out, phys)
        if cpus:
            info.update(gpu_node=node, cpus=cpus, source="nvidia-smi topo")
    except Exception:
        pass
    return info


def bind_process_near_gpu(device, report=False):
    """Pin this process (and the pinned host memory it allocates from now on: first touch) to the CPUs next to the
    GPU: with one process per GPU and no binding, the staging buffers of half the ranks end up on the other socket
    and every H2D/D2H crosses the inter-socket link.  Returns the NUMA node, or None when it cannot be determined or
    there is nothing to choose (a single node, no permission) -- then nothing is changed.  report=True returns the
    whole finding as a dict instead."""
    import os
    info = gpu_numa_topology(device)
    bound = False
    try:
        cpus = info["cpus"] & os.sched_getaffinity(0)
        if cpus and info["nodes"] != 1 and cpus != os.sched_getaffinity(0):
            os.sched_setaffinity(0, cpus)
            bound = True
    except Exception:
        pass
    if report:
        return {"nodes": info["nodes"], "gpu_node": info["gpu_node"], "source": info["source"], "bound": bound,
                "cpus_near_gpu": len(info["cpus"]), "cpus_allowed": len(os.sched_getaffinity(0))}
    return info["gpu_node"] if bound else None


def gather_slabs(black, white, rank, world, group=None, root=0, device=None):
    """Stack every rank's (Y, X/64) uint32 colour planes on rank `root`, slab r = rows [r*Y, (r+1)*Y) as in the
    reference's -d decomposition (optimized/main.cu:1542-1545): returns the two (world*Y, X/64) arrays there and
    (None, None) on the other ranks.  Backend-agnostic like HaloExchange: host tensors over gloo, staged through
    `device` memory over NCCL.  Every rank of the group must call it."""
    import numpy as np
    if world == 1:
        return np.ascontiguousarray(black), np.ascontiguousarray(white)
    out = []
    for plane in (black, white):
        t = torch.from_numpy(np.ascontiguousarray(plane, dtype=np.uint32).view(np.int32))
        if device is not None:
            t = t.to(device)
        parts = [torch.empty_like(t) for _ in range(world)] if rank == root else None
        dist.gather(t, parts, dst=root if group is None else dist.get_global_rank(group, root), group=group)
        out.append(torch.cat(parts).cpu().numpy().view(np.uint32) if rank == root else None)
    return out[0], out[1]


class HaloExchange:
    def __init__(self, rank, world, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.up = (rank - 1) % world      # owns the rows above mine
        self.down = (rank + 1) % world    # owns the rows below mine

    def start(self, backend, color):
        """Post the 4 messages of one colour.  Returns the request list (wait() before the ghosts
        are read).  Message order is fixed so that with world == 2 (up == down) the two rows
        exchanged with the single peer are matched unambiguously: first rows go first."""
        if self.world == 1:
            return []
        first, last = backend.edge_rows(color)
        top, bottom = backend.ghost_rows(color)
        ops = [dist.P2POp(dist.isend, first, self.up, self.group),      # my row 0   -> their bottom ghost
               dist.P2POp(dist.isend, last, self.down, self.group),     # my row Y-1 -> their top ghost
               dist.P2POp(dist.irecv, bottom, self.down, self.group),   # their row 0   -> my bottom ghost
               dist.P2POp(dist.irecv, top, self.up, self.group)]        # their row Y-1 -> my top ghost
        return dist.batch_isend_irecv(ops)

    @staticmethod
    def finish(reqs):
        for r in reqs:
            r.wait()


class _DevRows:
    """Zero-copy torch view of one device row of the library's lattice."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<i4", "data": (ptr, False),
                                         "version": 2}


class GpuSlab:
    """Backend over libising_b200.so for one rank (one GPU)."""

    def __init__(self, X, Y, rank, world, device, temp, seed, stream):
        self.eng = IsingEngine(X, Y, ndev=1, seed=seed, temp=temp, first_device=device, slab_first=rank,
                               slab_count=world, flags=F_NO_PEER_COPY if world > 1 else 0)
        self.eng.set_stream(0, stream.cuda_stream)
        self.Y = Y
        self._rows = {}
        if world > 1:
            dev = torch.device("cuda", device)
            for color in (0, 1):
                for row in (-1, 0, Y - 1, Y):
                    ptr, n = self.eng.row_ptr(0, color, row)
                    self._rows[(color, row)] = torch.as_tensor(_DevRows(ptr, n), device=dev)

    def edge_rows(self, color):
        return self._rows[(color, 0)], self._rows[(color, self.Y - 1)]

    def ghost_rows(self, color):
        return self._rows[(color, -1)], self._rows[(color, self.Y)]

    def update(self, color, it, part=PART_ALL, stream=None):
        self.eng.update_color_part(0, color, it, part, stream.cuda_stream if stream is not None else None)


class SlabEngine:
    """What bench.py and the multi-GPU tests drive: init + sweeps with overlapped halo exchange."""

    def __init__(self, X, Y, temp, seed, rank=0, world=1, device=0, group=None, overlap=True, halo="nccl"):
        """halo = "nccl": boundary rows travel as NCCL send/recv (overlapped with the interior when
        `overlap`); "fused": the update kernel stores them into the neighbours' ghost rows itself
        through CUDA-IPC-mapped peer memory (ising_ipc_attach), no per-colour communication call."""
        self.rank, self.world, self.overlap = rank, world, overlap
        self.X, self.Y, self.temp, self.seed, self.device = X, Y, temp, seed, device
        self.fused = (halo == "fused") and world > 1
        torch.cuda.set_device(device)
        self.main = torch.cuda.current_stream()
        self.comm = torch.cuda.Stream() if world > 1 else None
        self.slab = GpuSlab(X, Y, rank, world, device, temp, seed, self.main)
        self.eng = self.slab.eng
        self.halo = HaloExchange(rank, world, group)
        self.it = 0
        self._ev_edges = torch.cuda.Event()
        self._ev_comm = torch.cuda.Event()

    def _exchange_blocking(self, color):
        HaloExchange.finish(self.halo.start(self.slab, color))

    def _attach_peers(self):
        mine = self.eng.ipc_export()
        allh = [None] * self.world
        dist.all_gather_object(allh, mine, group=self.halo.group)
        self.eng.ipc_attach(allh[self.halo.up], allh[self.halo.down])
        self._attached = True

    def refresh_ghosts(self):
        """Fill the ghost rows from the neighbours' current edge rows (after init / import)."""
        if self.world == 1:
            return
        if self.fused:
            torch.cuda.synchronize()
            dist.barrier(group=self.halo.group)         # everybody's rows are final
            for color in (0, 1):
                self.eng.push_edges(color)
            torch.cuda.synchronize()
            dist.barrier(group=self.halo.group)         # everybody's ghosts are filled
        else:
            for color in (0, 1):
                self._exchange_blocking(color)

    def init_lattice(self):
        self.eng.init_lattice()
        self.it = 0
        if self.fused and not getattr(self, "_attached", False):
            self._attach_peers()
        self.refresh_ghosts()

    def _half_sweep(self, color, it):
        if self.world == 1:
            self.slab.update(color, it, PART_ALL)
            return
        if not self.overlap:
            self.slab.update(color, it, PART_ALL)
            self._exchange_blocking(color)
            return
        # edges first, then ship them while the interior runs
        self.main.wait_event(self._ev_comm)            # ghosts of the other colour have landed
        self.slab.update(color, it, PART_EDGES)
        self._ev_edges.record(self.main)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self._ev_edges)
            HaloExchange.finish(self.halo.start(self.slab, color))
            self._ev_comm.record(self.comm)
        self.slab.update(color, it, PART_INTERIOR)

    def sweeps(self, n):
        if self.fused:                       # one kernel per colour, halo pushed from inside it
            self.eng._ck(self.eng._L.ising_sweeps(self.eng._h, self.it + 1, int(n)))
            self.it += int(n)
            return
        for _ in range(n):
            self.it += 1
            self._half_sweep(0, self.it)
            self._half_sweep(1, self.it)
        if self.world > 1 and self.overlap:
            self.main.wait_event(self._ev_comm)

    def white_ghost_rows(self):
        """The two white ghost rows (above, below the slab) as a (2, X/64) uint32 host array: what
        ising_run_slab_batch_from_host wants next to a slab uploaded from the host."""
        import numpy as np
        torch.cuda.synchronize()
        top, bottom = self.slab.ghost_rows(1)
        return np.stack([top.cpu().numpy(), bottom.cpu().numpy()]).view(np.uint32)

    def run_batch_from_host(self, other, black_in, white_in, ghosts_in, n, black_out=None, white_out=None):
        """Replica batch on this rank's slab (see IsingEngine.run_batch_from_host); `other` is a second
        SlabEngine of the same shape on the same GPU, or a list of them (jobs round-robin over [self] + other).
        Every rank of the group must make the same call."""
        oe = [o.eng for o in other] if isinstance(other, (list, tuple)) else (other.eng if other else None)
        if self.world == 1:
            return self.eng.run_batch_from_host(oe, black_in, white_in, n, black_out, white_out, first_it=self.it + 1)
        if not self.fused:
            raise ValueError("replica batches over several slabs need halo='fused'")
        return self.eng.run_batch_from_host(oe, black_in, white_in, n, black_out, white_out,
                                            first_it=self.it + 1, ghosts_in=ghosts_in)

    def count_spins(self):
        return self.eng.count_spins()

    # -- observables of the WHOLE lattice (every rank of the group must make the same call) -----------
    def bond_sum(self):
        """Sum of s_i*s_j over the nearest-neighbour bonds of all slabs (ising_bond_sum per slab -- its lower
        bonds read the white ghost rows, current after sweeps() -- added up over the ranks)."""
        s = self.eng.bond_sum()
        if self.world > 1:
            t = torch.tensor([s], dtype=torch.int64, device=torch.device("cuda", self.device))
            dist.all_reduce(t, group=self.halo.group)
            s = int(t.item())
        return s

    def energy_per_site(self):
        return -self.bond_sum() / float(self.world * self.Y * self.X)

    def correlation(self, ncorr=128, root=0):
        """Two-point function of the whole lattice, the reference's -c (optimized/main.cu:1072-1138 walks the slabs
        of all devices of its one process).  ising_correlation needs every slab in one context, so the slabs' bit
        planes (1 bit per spin) are gathered on rank `root`, loaded into a 1-device context of world*Y rows next to
        its slab and evaluated there; every rank gets the result.  Off the hot path: print cadence only."""
        if self.world == 1:
            return self.eng.correlation(ncorr)
        import numpy as np
        torch.cuda.synchronize()
        b, w = self.eng.export_bits()
        B, W = gather_slabs(b, w, self.rank, self.world, self.halo.group, root, torch.device("cuda", self.device))
        res = [None]
        if self.rank == root:
            with IsingEngine(self.X, self.Y * self.world, seed=self.seed, temp=self.temp, first_device=self.device) as whole:
                whole.import_bits(B, W)
                res[0] = np.asarray(whole.correlation(ncorr))
        dist.broadcast_object_list(res, src=root if self.halo.group is None else dist.get_global_rank(self.halo.group, root),
                                   group=self.halo.group)
        return res[0]

    def close(self):
        self.eng.close()

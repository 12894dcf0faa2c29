#!/usr/bin/env python
"""bench.py -- spin-flip attempts/ns of the checkerboard Metropolis hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json): 65536 x 65536 lattice per GPU at T = 1.5, default seed; at N GPUs the
lattice is N slabs stacked along Y (131072 x 65536 at 2, 262144 x 65536 at 4 and -- as the
north star names it -- 262144 x 65536 at 8, i.e. 32768 rows per GPU there; --rows-per-gpu
overrides).  A "step" is --sweeps-per-step (16) full sweeps (black + white colour update with
its halo exchange).  The timed region holds exactly K steps, bracketed by a barrier and a
device synchronize, timed with CUDA events on the launching stream, max over ranks.

One JSON line on rank 0.  `value` = total flip attempts of all ranks / that time, lattice
resident in HBM.  `e2e` = same metric through the C ABI with HOST buffers, three ways: `value` =
replica batch (every step uploads both colour planes from pinned memory, sweeps, downloads the
step's observable), `with_lattice_download` = the same with the final planes downloaded too,
`dependent_chain` = step k+1 starts from the planes step k downloaded (nothing can overlap).
`roofline` = algorithmic bytes (0.25 B per flip attempt) over the measured kernel
time against the measured HBM peak; `cpu_baseline` = numpy restatement of basic_python on
BASELINE config 1 (256 x 256, T = 2.5, 128 sweeps; 1 core), `cpu_baseline_large` the same code at
4096 x 4096, `cpu_oracle` the C/OpenMP oracle on all cores.

Before the timed region both arms start from a fresh lattice and run steps x sweeps-per-step
sweeps; the spin counts at that point are printed as `counts_at_sweeps` -- equal numbers in the
two arms' lines prove bit-equality of the whole lattice history on the driver's own box (the counts
at sweeps 16..128 are also asserted against the README transcripts where the lattice is one of
theirs: `kat`).  `variants.tensorcore` = BASELINE config 4 (32768 x 32768 tensor-core variant,
updates/ns) measured the same way in both arms.  At N > 1 `selfcheck` compares a small slab
lattice run by all ranks with the 1-GPU engine before anything is timed.

--impl reference: runs the UNMODIFIED reference (optimized/main.cu rebuilt for sm_100a as
oracle/_ref/cuIsing_ref, tensorcore/main.cu as oracle/_ref/ising_tensorcore_ref) through its own
CLI on the same lattice; under torchrun only rank 0 runs it (cuIsing_ref -d N drives all N GPUs
from one process, as the reference does).  If the binary is absent it times the CPU oracle port on
all host cores instead and says so.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_GUARD = None
X_COLS = 65536
TEMP = 1.5
ALG_BYTES_PER_FLIP = 0.25      # each spin read once + written once per sweep at 1 bit/spin
REF_BYTES_PER_FLIP = 1.5       # the reference's own accounting, optimized/main.cu:1887-1890


# tensor-core variant, BASELINE config 4 / SURVEY 8(d): 32768 x 32768 (-x 128 -y 128 in 256-blocks),
# -a 0.661030 (T = 1.5), seed 1234; warm-up and trial iterations of the bench leg (reference CLI -w / -n)
TC_NB, TC_ALPHA, TC_SEED, TC_WARM, TC_ITERS = 128, 0.661030, 1234, 20, 200
TC_ALG_BYTES_PER_UPDATE = 2.0  # int8 lattice: every spin read once and written once per sweep

# optimized/README.md:128-137 (65536 x 65536) and :240-249 (131072 x 65536): up-spin counts after n sweeps
README_KAT = {
    (1, 65536): ("optimized/README.md:128-137", {16: 2147575418, 32: 2147641872, 48: 2147605659, 64: 2147701147,
                                                   80: 2147558546, 96: 2147471275, 112: 2147612509, 128: 2147678887}),
    (2, 65536): ("optimized/README.md:240-249", {16: 4294617248, 32: 4293898346, 48: 4292806461, 64: 4291852263,
                                                   80: 4291086016, 96: 4290256223, 112: 4289621029, 128: 4288877118}),
    (8, 65536): ("optimized/README.md:307-316", {16: 17176389528, 32: 17172963073, 48: 17170610910, 64: 17168843228,
                                                   80: 17167009008, 96: 17165014291, 112: 17163708078, 128: 17162287230}),
}
DEFAULT_SEED = 463463564571


def default_rows(n):
    return 32768 if n >= 8 else 65536


def workload_config(rows, n, seed, S):
    """`config` of the JSON line: identical in both arms for the same command line."""
    return {"workload": "checkerboard Metropolis, %dx%d lattice (%d rows per GPU), T=%.1f, seed %d, "
                        "%d sweeps per step" % (rows * n, X_COLS, rows, TEMP, seed, S),
            "l2_flush": "inputs larger than L2 (>= %d MiB of spins per GPU vs 126 MB L2)" % (rows * X_COLS // 8 >> 20),
            "variants": "tensorcore: %dx%d tensor-core variant, alpha %.6f, seed %d, %d warm-up + %d trial iterations "
                        "(1 GPU)" % (256 * TC_NB, 256 * TC_NB, TC_ALPHA, TC_SEED, TC_WARM, TC_ITERS)}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch():
    """dram bytes per update launch from the committed ncu capture, or None."""
    p = os.path.join(ROOT, "profiles", "update_kernel_ncu_summary.json")
    try:
        d = json.load(open(p))
        return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
    except Exception:
        return None


def ncu_pipe_summary():
    """What actually binds the kernel (committed ncu capture): busiest pipe and issue-slot use, or None."""
    p = os.path.join(ROOT, "profiles", "update_kernel_ncu_summary.json")
    try:
        d = json.load(open(p))
        return {"fmaheavy_pipe_pct": round(float(d["pipe_fmaheavy_pct"]), 1), "alu_pipe_pct": round(float(d["pipe_alu_pct"]), 1),
                "issue_slots_pct": round(float(d["issue_active_pct"]), 1), "source": "profiles/update_kernel_ncu_summary.json"}
    except Exception:
        return None


def cpu_baselines():
    from oracle import ising_oracle, numpy_basic
    # BASELINE config 1: basic_python on 256 x 256, T = 2.5, 128 sweeps (repeated to fill ~10 s of CPU work)
    reps, secs, fl, mag = 0, 0.0, [], 0.0
    while secs < 10.0 and reps < 64:
        f, mag, t = numpy_basic.run(n=256, m=256, alpha=2.5 / numpy_basic.TCRIT, sweeps=128, seed=1234 + reps)
        fl.append(f); secs += t; reps += 1
    base = {"value": round(sorted(fl)[len(fl) // 2], 5), "unit": "flips/ns", "cores": 1, "host_cores": os.cpu_count(),
            "kind": "port", "abs_magnetisation_last_run": round(mag, 5),
            "sample": "numpy restatement of basic_python/ising_basic.py update_lattice (pinned against the reference "
                      "kernel under numba's simulator, tests/test_numpy_baseline.py), BASELINE config 1: 256x256, T=2.5, "
                      "128 sweeps, median of %d runs, %.1f s" % (reps, secs)}
    n = 4096
    sweeps = 12
    fl, mag, secs = numpy_basic.run(n=n, m=n, alpha=TEMP / numpy_basic.TCRIT, sweeps=sweeps, seed=1234)
    large = {"value": round(fl, 5), "unit": "flips/ns", "cores": 1, "host_cores": os.cpu_count(), "kind": "port",
             "sample": "same code, %dx%d, T=%.1f, %d sweeps, %.1f s" % (n, n, TEMP, sweeps, secs)}
    L = ising_oracle.OracleLattice(X_COLS, 4096)
    L.init()
    t0 = time.perf_counter()
    L.sweeps(2, TEMP)
    dt = time.perf_counter() - t0
    L.close()
    orc = {"value": round(X_COLS * 4096 * 2 / dt * 1e-9, 4), "unit": "flips/ns", "cores": ising_oracle.num_threads(),
           "kind": "port", "sample": "bit-exact C/OpenMP oracle of optimized/, 4096 rows x 65536, 2 sweeps, %.1f s" % dt}
    return base, large, orc


def tensorcore_leg(device):
    """BASELINE config 4 through the C ABI (ising_tc_*): 32768 x 32768, warm-up + trial iterations numbered as the
    reference numbers them; CUDA events on the launching stream inside ising_tc_update_timed."""
    from ising_gpu_b200.tensorcore import TensorcoreEngine
    peak, peak_src = measured_peaks()
    traffic = None
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "tc_pipe_kernel_ncu_summary.json")))
        traffic = float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
    except Exception:
        pass
    sampler = ClockSampler(device)
    with TensorcoreEngine(TC_NB, TC_NB, seed=TC_SEED, alpha=TC_ALPHA, device=device) as T:
        T.init()
        T.update(TC_WARM)
        l0 = T.launches
        ms = T.update_timed(TC_ITERS)
        launches = T.launches - l0
        n = float(T.nx) * T.ny
        total = T.sum()
    clocks = sampler.stop()
    upd = n * TC_ITERS / (ms * 1e6)
    kern_ms = ms / (2 * TC_ITERS)          # one colour update = one launch of the persistent kernel
    alg = n / 2 * TC_ALG_BYTES_PER_UPDATE
    ach = alg / (kern_ms * 1e-3) / 1e9
    return {"metric": "updates/ns", "value": round(upd, 2), "unit": "updates/ns", "iterations": TC_ITERS,
            "warmup_iterations": TC_WARM, "ms_per_iteration": round(ms / TC_ITERS, 4), "dtype": "int8 (tcgen05.mma kind::i8)",
            "lattice": "%dx%d" % (T.nx, T.ny), "gpu_launches": int(launches),
            "magnetism_after": {"iterations": TC_WARM + TC_ITERS, "value": "%g" % (abs(total) / n), "sum": total},
            "roofline": {"bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                         "traffic": traffic, "kernel": "tc_update_pipe_kernel", "kernel_ms": round(kern_ms, 4),
                         "algorithmic_bytes_per_launch": alg, "peak_source": peak_src,
                         "tensor_ops_per_update": 512},
            "clocks": clocks}


def selfcheck_slabs(a, rank, world, local):
    """N > 1: a small lattice cut into `world` slabs and run by all ranks (same halo transport as the bench)
    must equal the 1-GPU engine on the stacked lattice, bit for bit, after 3 sweeps."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from ising_gpu_b200.engine import IsingEngine
    from ising_gpu_b200.slab import SlabEngine
    X, Y, n = 4096, 64, 3
    with torch.cuda.stream(torch.cuda.Stream()):
        se = SlabEngine(X, Y, TEMP, a.seed, rank=rank, world=world, device=local, overlap=not a.no_overlap, halo=a.halo)
    se.init_lattice()
    se.sweeps(n)
    torch.cuda.synchronize()
    b, w = se.eng.export_bits()
    allb = [None] * world
    dist.all_gather_object(allb, (b, w))
    se.close()
    ok = True
    if rank == 0:
        with IsingEngine(X, Y * world, seed=a.seed, temp=TEMP, first_device=local) as one:
            one.init_lattice()
            one.sweeps(n)
            ob, ow = one.export_bits()
        ok = bool(np.array_equal(ob, np.vstack([x[0] for x in allb])) and np.array_equal(ow, np.vstack([x[1] for x in allb])))
    t = torch.tensor([float(ok)], device="cuda", dtype=torch.float64)
    dist.broadcast(t, 0)
    return {"ok": bool(t.item()), "lattice": "%dx%d in %d slabs of %d rows, halo=%s" % (Y * world, X, world, Y, a.halo),
            "sweeps": n, "against": "1-GPU engine on the stacked lattice, all spins compared"}


def run_b200(a):
    import torch
    import torch.distributed as dist
    import numpy as np
    from ising_gpu_b200.slab import SlabEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (a.gpus, world))
    torch.cuda.set_device(local)
    numa = None
    if world > 1:
        from ising_gpu_b200.slab import bind_process_near_gpu
        numa = bind_process_near_gpu(local, report=True)     # host staging buffers next to the GPU's PCIe root
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rows = a.rows_per_gpu or default_rows(world)
    S = a.sweeps_per_step

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    selfcheck = selfcheck_slabs(a, rank, world, local) if world > 1 and not a.no_selfcheck else None
    if selfcheck is not None and not selfcheck["ok"]:
        raise SystemExit("selfcheck failed: the slab decomposition does not reproduce the 1-GPU lattice: %r" % (selfcheck,))

    work = torch.cuda.Stream() if a.own_stream else torch.cuda.current_stream()
    with torch.cuda.stream(work):
        eng = SlabEngine(X_COLS, rows, TEMP, a.seed, rank=rank, world=world, device=local, overlap=not a.no_overlap,
                         halo=a.halo)

    def total_counts():
        up, dn = eng.count_spins()
        if world > 1:
            t = torch.tensor([up, dn], device="cuda", dtype=torch.int64)
            dist.all_reduce(t)
            up, dn = int(t[0]), int(t[1])
        return up, dn

    # ---- fresh lattice, steps x S sweeps: the counts both arms print (and, where the lattice is one of the
    # README's, every 16th sweep up to 128 against its transcript).  Doubles as extra warm-up.
    eng.init_lattice()
    kat_src, kat_tab = README_KAT.get((world, rows), (None, {})) if (a.seed == DEFAULT_SEED and TEMP == 1.5) else (None, {})
    kat_checked, kat_ok, done = [], True, 0
    total = S * a.steps
    while done < total:
        nxt = min(total, (done // 16 + 1) * 16) if kat_tab and done < 128 else total
        eng.sweeps(nxt - done)
        done = nxt
        if done in kat_tab:
            up, _ = total_counts()
            kat_checked.append(done)
            kat_ok = kat_ok and (up == kat_tab[done])
    cas_up, cas_dn = total_counts()
    counts_at = {"sweeps": total, "up": cas_up, "down": cas_dn, "from": "fresh lattice (Philox init), it = 1.."}
    kat = {"source": kat_src, "sweeps_checked": kat_checked, "ok": kat_ok} if kat_checked else None
    if kat is not None and not kat_ok:
        raise SystemExit("spin counts differ from the reference's README transcript (%s)" % kat_src)

    for _ in range(a.warmup):
        eng.sweeps(S)
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = eng.eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(work)                     # events on the stream the kernels are launched on
    for _ in range(a.steps):
        eng.sweeps(S)
    e1.record(work)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.eng.launches - launches0
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    up, dn = total_counts()
    flips_per_rank = float(X_COLS) * rows * S * a.steps
    value = flips_per_rank * world / (ms * 1e6)

    # ---- end to end through the C ABI with host buffers ------------------------------------
    e2e = None
    if not a.no_e2e:
        pitch = X_COLS // 64
        plane_bytes = rows * pitch * 4
        ksteps = max(1, min(a.steps, a.e2e_steps)) if a.e2e_steps else a.steps
        flips_e2e = float(X_COLS) * rows * S * ksteps * world

        def pinned(wc=False):
            if wc:      # write-combined input staging (ising_host_alloc): the host only ever writes these
                from ising_gpu_b200.engine import HostBuffer
                hb = HostBuffer((rows, pitch), write_combined=True)
                return hb, hb.array
            t = torch.empty((rows, pitch), dtype=torch.int32).pin_memory()
            return t, t.numpy().view(np.uint32)

        def allmax(*vals):
            if world == 1:
                return list(vals)
            t = torch.tensor(list(vals), device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return [float(x) for x in t.tolist()]
        b0, w0 = eng.eng.export_bits()

        def chain_e2e(nb, nw):
            """Dependent chain: step k+1 starts from the planes step k downloaded -- what a caller whose state
            lives in host memory between steps pays; nothing can overlap."""
            def e2e_step():
                eng.eng.import_bits(nb, nw)          # H2D of both colour planes (pinned)
                eng.refresh_ghosts()                  # ghost rows belong to the uploaded state
                eng.sweeps(S)
                eng.eng._ck(eng.eng._L.ising_export_bits(eng.eng._h, 0, nb.ctypes.data, nw.ctypes.data))  # D2H
                return eng.count_spins()              # D2H of the step's observable
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(ksteps):
                e2e_step()
            barrier()
            return time.perf_counter() - t0

        keep, ins, outs = [], [], []
        NCTX = max(1, a.e2e_contexts)   # contexts per GPU the jobs rotate over
        ready = True
        try:        # pinned host memory is the one local resource that can run out here
            for j in range(2):          # inputs: two lattices, uploaded alternately
                (tb, nb), (tw, nw) = pinned(a.staging == "wc"), pinned(a.staging == "wc")
                nb[:], nw[:] = b0, w0
                keep += [tb, tw]
                ins.append((nb, nw))
            for j in range(NCTX):       # outputs: one landing zone per context in flight
                (tob, nob), (tow, now) = pinned(), pinned()
                keep += [tob, tow]
                outs.append((nob, now))
        except (RuntimeError, Exception):
            ready = False
        ready = bool(-allmax(-float(ready))[0])           # every rank enters the batch, or none does
        runs_dl, runs = [], []
        batch_error = None
        e2e_clocks = {}
        ok = False
        api = None
        if ready and world == 1:
            # replica batch: every step is one job of the batch = upload both planes of a lattice from
            # pinned memory, S sweeps, spin counts, download both planes; jobs alternate between two
            # contexts so the copies of one step overlap the kernels of the next (ising_run_batch_from_host)
            from ising_gpu_b200.engine import IsingEngine
            others = [IsingEngine(X_COLS, rows, seed=a.seed, temp=TEMP, first_device=local) for _ in range(NCTX - 1)]
            # one unpipelined step first: the result every job of the batch must reproduce
            it0 = eng.it
            eng.eng.import_bits(b0, w0)
            eng.sweeps(S)
            want_b, want_w = eng.eng.export_bits()
            want_counts = eng.count_spins()
            eng.it = it0

            def batch(k, download):
                return eng.eng.run_batch_from_host(others, [ins[j & 1][0] for j in range(k)],
                                                   [ins[j & 1][1] for j in range(k)], S,
                                                   [outs[j % NCTX][0] for j in range(k)] if download else None,
                                                   [outs[j % NCTX][1] for j in range(k)] if download else None,
                                                   first_it=eng.it + 1)
            api = ("ising_run_batch_from_host_n (C ABI, pinned host buffers): per step H2D of both planes, %d sweeps, "
                   "count kernels, D2H of the spin counts; %d contexts, copies overlap the other steps' kernels; result "
                   "checked against the unpipelined path" % (S, NCTX))
        elif ready and a.halo == "fused":
            # the same replica batch on every rank's slab (ising_run_slab_batch_from_host): per step H2D of the
            # slab's planes and white ghost rows, S sweeps with the fused halo push between the ranks, counts,
            # D2H of the planes; two contexts per rank so that copies overlap the other step's kernels
            others = []
            for _ in range(NCTX - 1):
                with torch.cuda.stream(torch.cuda.Stream()):   # SlabEngine works on the stream current at construction
                    others.append(SlabEngine(X_COLS, rows, TEMP, a.seed, rank=rank, world=world, device=local, halo="fused"))
                others[-1].init_lattice()
            # one unpipelined step first: gives the ghost rows that belong to the uploaded state and the
            # result the pipelined steps must reproduce
            it0 = eng.it
            eng.eng.import_bits(b0, w0)
            eng.refresh_ghosts()
            ghosts = eng.white_ghost_rows()
            eng.sweeps(S)
            want_b, want_w = eng.eng.export_bits()
            want_counts = eng.count_spins()
            eng.it = it0
            for o in others:
                o.it = it0
            tg = torch.empty((2, pitch), dtype=torch.int32).pin_memory()
            ng = tg.numpy().view(np.uint32)
            ng[:] = ghosts

            def batch(k, download):
                return eng.run_batch_from_host(others, [ins[j & 1][0] for j in range(k)],
                                               [ins[j & 1][1] for j in range(k)], [ng] * k, S,
                                               [outs[j % NCTX][0] for j in range(k)] if download else None,
                                               [outs[j % NCTX][1] for j in range(k)] if download else None)
            api = ("ising_run_batch_from_host_n (C ABI, pinned host buffers): per step and rank H2D of the slab's planes "
                   "and white ghost rows, %d sweeps with the fused halo push, count kernels, D2H of the spin counts; "
                   "%d contexts per rank, copies overlap the other steps' kernels; result checked against the "
                   "unpipelined path" % (S, NCTX))
        else:
            batch = None
        if batch is not None:
            def timed(download):
                batch(max(NCTX, 8), download)        # untimed: every context touched, clocks back up after host-side checks
                out = []
                smp = ClockSampler(local) if rank == 0 else None
                for _ in range(5):                   # five runs of K steps, all reported; the best one is the value (single
                                                     # runs are hit by host-side hiccups: 1050..1580 at N=1, same clocks)
                    barrier()                        # (the reference arm does the same)
                    t0 = time.perf_counter()
                    counts = batch(ksteps, download)
                    barrier()
                    out.append(time.perf_counter() - t0)
                if smp:
                    e2e_clocks["with_lattice_download" if download else "counts_only"] = smp.stop()
                return out, counts
            runs_dl, counts = timed(True)            # D2H of the final planes as well (a checkpoint per step)
            ok = all(c == want_counts for c in counts) and all(
                np.array_equal(ob, want_b) and np.array_equal(ow, want_w) for ob, ow in outs[:min(NCTX, ksteps)])
            runs, counts2 = timed(False)             # D2H of the step's observable only
            ok = ok and counts2 == counts
            ok = bool(-allmax(-float(ok))[0])
            runs_dl = [allmax(x)[0] for x in runs_dl]
            runs = [allmax(x)[0] for x in runs]
            for o in others:
                o.close()
            batch_error = None
            if not ok:   # never report a number for a wrong result: the line falls back to the dependent chain and says so
                batch_error = "pipelined replica batch did not reproduce the unpipelined result; its timings are withheld"
                batch = None
        # dependent chain on one of the input buffers (always measured; the only e2e number without a batch)
        if outs:     # (plain page-locked buffers: the chain downloads into what it uploads next)
            outs[0][0][:], outs[0][1][:] = b0, w0
            cb, cw = outs[0]
        else:
            (tb, cb), (tw, cw) = pinned(), pinned()
            cb[:], cw[:] = b0, w0
        del b0, w0
        dt_chain = allmax(chain_e2e(cb, cw))[0]
        chain = {"value": round(flips_e2e / (dt_chain * 1e9), 2), "unit": "flips/ns",
                 "h2d_bytes_per_step": 2 * plane_bytes * world, "d2h_bytes_per_step": (2 * plane_bytes + 8) * world,
                 "api": "ising_import_bits + ising_sweeps + ising_export_bits + ising_count_spins per step (C ABI, pinned "
                        "host buffers): step k+1 uploads what step k downloaded, nothing overlaps"}
        if batch is not None:
            h2d = (2 * plane_bytes + (2 * pitch * 4 if world > 1 else 0)) * world
            e2e = {"value": round(flips_e2e / (min(runs) * 1e9), 2), "unit": "flips/ns",
                   "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8 * world, "steps": ksteps,
                   "timer": "host monotonic clock around the synchronised region, max over ranks; five runs of %d steps, "
                            "the best one is `value`" % ksteps,
                   "runs": [round(flips_e2e / (x * 1e9), 2) for x in runs], "api": api, "numa_of_rank0": numa, "input_staging": a.staging, "contexts_per_gpu": NCTX,
                   "clocks": e2e_clocks,
                   "with_lattice_download": {"value": round(flips_e2e / (min(runs_dl) * 1e9), 2), "unit": "flips/ns",
                                             "runs": [round(flips_e2e / (x * 1e9), 2) for x in runs_dl],
                                             "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": (2 * plane_bytes + 8) * world},
                   "dependent_chain": chain}
        else:
            e2e = dict(chain)
            e2e.update({"steps": ksteps, "timer": "host monotonic clock around the synchronised region, max over ranks",
                        "numa_of_rank0": numa, "dependent_chain": chain,
                        "note": batch_error or "no replica batch on this transport (halo=%s%s)" % (
                            a.halo, "" if ready else ", pinned host memory exhausted")})

    out = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        half_sweeps = 2 * S * a.steps
        kern_ms = ms / half_sweeps                         # one colour half-sweep on one GPU
        alg_bytes = float(X_COLS) * rows / 2 * ALG_BYTES_PER_FLIP
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": ncu_traffic_per_launch(),
                "kernel": eng.eng.kernel_name, "kernel_ms": round(kern_ms, 4),
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "reference_accounting_gbs": round(achieved * REF_BYTES_PER_FLIP / ALG_BYTES_PER_FLIP, 1),
                "binding_resource": ncu_pipe_summary(),
                "note": "HBM-nominal path, but the bit-exact Philox4x32-10 per site makes it bound by integer issue "
                        "(16 IMAD.WIDE + 18 LOP3 + 8 ISETP per Philox call; FMA-heavy pipe 87 % busy, the ALU/dispatch "
                        "side binding): DESIGN.md 4.1 and profiles/update_kernel_ncu_summary.json"}
        out = {"metric": "spin-flip attempts/ns", "value": round(value, 2), "unit": "flips/ns", "n_gpus": world,
               "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms / a.steps, 4), "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
               "config": workload_config(rows, world, a.seed, S),
               "details": {"lattice_bytes_per_gpu": rows * X_COLS // 8,
                           "halo": "none" if world == 1 else (
                               "fused: update kernel stores its edge rows into the neighbours' ghost rows (CUDA IPC "
                               "peer memory over NVLink) and signals device-side counters" if a.halo == "fused" else
                               "NCCL send/recv of 1 row per direction per colour, " +
                               ("overlapped with the interior update" if not a.no_overlap else "serialised")),
                           "update_kernel": os.environ.get("ISING_B200_KERNEL", "default")},
               "roofline": roof, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
               "counts_at_sweeps": counts_at, "kat": kat, "selfcheck": selfcheck,
               "final_counts": {"up": up, "down": dn, "sweeps": total + S * (a.warmup + a.steps)}}
    if world > 1:
        dist.barrier()
    eng.close()
    if rank == 0:
        variants = {}
        if world == 1 and not a.no_tc:
            variants["tensorcore"] = tensorcore_leg(local)
        out["variants"] = variants
        if world == 1 and not a.no_cpu:
            out["cpu_baseline"], out["cpu_baseline_large"], out["cpu_oracle"] = cpu_baselines()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        _GUARD.emit(json.dumps(out))


def reference_tensorcore_leg():
    """The reference's tensorcore/ program (oracle/_ref/ising_tensorcore_ref, stock CLI) on BASELINE config 4;
    "updates per ns" and "average magnetism" parsed from its REPORT block (tensorcore/main.cu:601-609, :638)."""
    binary = os.path.join(ROOT, "oracle", "_ref", "ising_tensorcore_ref")
    if not os.access(binary, os.X_OK):
        return {"unavailable": "oracle/_ref/ising_tensorcore_ref not built"}
    cmd = [binary, "-x", str(TC_NB), "-y", str(TC_NB), "-w", str(TC_WARM), "-n", str(TC_ITERS), "-a", "%.6f" % TC_ALPHA,
           "-s", str(TC_SEED)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    except subprocess.TimeoutExpired:
        return {"unavailable": "reference tensorcore run timed out"}
    m = re.search(r"updates per ns:\s*(\S+)", r.stdout)
    if not m:
        return {"unavailable": "reference tensorcore run failed: " + (r.stderr or r.stdout)[-300:]}
    el = re.search(r"elapsed time:\s*(\S+) sec", r.stdout)
    mg = re.search(r"average magnetism \(absolute\):\s*(\S+)", r.stdout)
    return {"metric": "updates/ns", "value": float(m.group(1)), "unit": "updates/ns", "iterations": TC_ITERS,
            "warmup_iterations": TC_WARM, "ms_per_iteration": round(float(el.group(1)) * 1e3 / TC_ITERS, 4) if el else None,
            "dtype": "fp16 (cuBLAS batched HGEMM, tensor op)", "lattice": "%dx%d" % (256 * TC_NB, 256 * TC_NB),
            "magnetism_after": {"iterations": TC_WARM + TC_ITERS, "value": mg.group(1) if mg else None},
            "reference_kind": "tensorcore/main.cu for sm_100a, only its final CUB reduce replaced "
                              "(oracle/build_ref_tensorcore.py), run through its own CLI: " + " ".join(cmd[1:])}


def run_reference(a):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = max(world, a.gpus)
    rows = a.rows_per_gpu or default_rows(n)
    S = a.sweeps_per_step
    binary = os.path.join(ROOT, "oracle", "_ref", "cuIsing_ref")
    base = {"metric": "spin-flip attempts/ns", "unit": "flips/ns", "n_gpus": n, "steps": a.steps, "warmup": a.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic", "impl": "reference",
            "config": workload_config(rows, n, a.seed, S)}
    try:
        have_gpu = subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0
    except OSError:
        have_gpu = False
    if os.access(binary, os.X_OK) and have_gpu and not a.ref_cpu:
        # the reference has no warm-up option: run the warm-up sweeps as a separate short run
        def call(nsweeps):
            cmd = [binary, "-x", str(X_COLS), "-y", str(rows), "-n", str(nsweeps), "-t", str(TEMP), "-s", str(a.seed),
                   "-d", str(n)]
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
            m = re.search(r"Kernel execution time for (\d+) update steps: (\S+) ms, (\S+) flips/ns", r.stdout)
            if not m:
                raise RuntimeError("reference run failed: " + (r.stderr or r.stdout)[-400:])
            fin = re.search(r"Final\s+magnetization:\s+\S+, up_s:\s+(\d+), dw_s:\s+(\d+)", r.stdout)
            return float(m.group(2)), float(m.group(3)), (int(fin.group(1)), int(fin.group(2))) if fin else None
        try:
            if a.warmup:
                call(S * a.warmup)      # same number of warm-up sweeps as the other arm (separate process: the
                                        # reference CLI has no warm-up option; clocks and driver state are warm)
            runs = [call(S * a.steps) for _ in range(5)]    # five runs, all reported, the best one is the value
            ms, fl, fin = max(runs, key=lambda r: r[1])
            base.update({"value": fl, "ms_per_step": round(ms / a.steps, 4), "dtype": "u64 (4 bit/spin)",
                         "runs": [r[1] for r in runs],
                         "reference_kind": "unmodified optimized/main.cu + utils.c, nvcc -O3 --use_fast_math for "
                                           "sm_100a, run through its own CLI (cuIsing_ref -d %d), lattice resident "
                                           "on the GPU(s)" % n,
                         "counts_at_sweeps": {"sweeps": S * a.steps, "up": fin[0], "down": fin[1],
                                              "from": "fresh lattice (its own init), its Final magnetization line"} if fin else None,
                         "final_counts": {"up": fin[0], "down": fin[1], "sweeps": S * a.steps} if fin else None,
                         "cpu_baseline": {"value": fl, "unit": "flips/ns", "cores": 0, "kind": "reference",
                                          "sample": "GPU reference binary, not a CPU run: %d sweeps of the full "
                                                    "workload" % (S * a.steps)},
                         "e2e": {"value": fl, "unit": "flips/ns", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                                 "note": "resident only: the reference CLI has no host-buffer entry point, its lattice "
                                         "never leaves the GPU"}})
            base["variants"] = {"tensorcore": reference_tensorcore_leg()} if (n == 1 and not a.no_tc) else {}
            _GUARD.emit(json.dumps(base))
            return
        except Exception as e:   # fall through to the CPU port
            base["reference_gpu_error"] = str(e)[:300]
    # CPU fallback: the bit-exact oracle port on all host cores, bounded sample of the workload
    from oracle import ising_oracle
    srows = 2048
    L = ising_oracle.OracleLattice(X_COLS, srows)
    L.init()
    for _ in range(min(a.warmup, 1)):
        L.sweeps(1, TEMP)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        L.sweeps(1, TEMP)
    dt = time.perf_counter() - t0
    fl = X_COLS * srows * a.steps / dt * 1e-9
    base.update({"value": round(fl, 4), "ms_per_step": round(dt * 1e3 / a.steps, 3), "dtype": "u8",
                 "cpu_baseline": {"value": round(fl, 4), "unit": "flips/ns", "cores": ising_oracle.num_threads(),
                                  "kind": "port", "sample": "C/OpenMP oracle, %d rows x %d, 1 sweep per step" %
                                  (srows, X_COLS)},
                 "e2e": {"value": round(fl, 4), "unit": "flips/ns", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    _GUARD.emit(json.dumps(base))


class _StdoutGuard:
    """Everything libraries print to fd 1 (e.g. NCCL's version banner) goes to stderr; only the
    final JSON line reaches the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.write(self.real, (line + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sweeps-per-step", type=int, default=16)
    ap.add_argument("--rows-per-gpu", type=int, default=0)
    ap.add_argument("--seed", type=int, default=463463564571)
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the e2e leg (0 = --steps)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-contexts", type=int, default=3,
                    help="contexts per GPU the e2e replica batch rotates over (2: a download and the next upload share "
                         "one job's compute time; 3: they have two)")
    ap.add_argument("--own-stream", type=int, default=0, help="1: the lattice works on its own non-blocking stream; "
                                                              "0: on torch's current (default) stream.  No measurable "
                                                              "difference, e2e hiccups included (profiles/r02_e2e_noise.txt)")
    ap.add_argument("--staging", default="pinned", choices=["pinned", "wc"],
                    help="e2e input buffers: plain page-locked memory, or write-combined (ising_host_alloc)")
    ap.add_argument("--no-tc", action="store_true", help="skip the tensor-core variant leg (variants.tensorcore)")
    ap.add_argument("--no-selfcheck", action="store_true", help="N > 1: skip the small slab-vs-1-GPU comparison")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--halo", default="fused", choices=["fused", "nccl"],
                    help="fused: the update kernel stores boundary rows into the neighbours' ghost rows over NVLink "
                         "(CUDA IPC peer memory); nccl: NCCL send/recv per colour")
    ap.add_argument("--ref-cpu", action="store_true", help="reference arm: force the CPU oracle port")
    a = ap.parse_args()
    global _GUARD
    _GUARD = _StdoutGuard()
    if a.warmup < 3 and a.impl == "b200":
        a.warmup = 3   # timing rule: at least 3 warm-up steps
    if a.impl == "reference":
        run_reference(a)
    else:
        if a.gpus > 1 and int(os.environ.get("WORLD_SIZE", "1")) == 1:
            raise SystemExit("for --gpus N>1 launch with torchrun (one rank per GPU)")
        run_b200(a)


if __name__ == "__main__":
    main()

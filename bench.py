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
resident in HBM.  `e2e` = same metric through the C ABI with HOST buffers: every step uploads
both colour planes from pinned memory, sweeps, downloads the planes and the spin counts.
`roofline` = algorithmic bytes (0.25 B per flip attempt) over the measured kernel
time against the measured HBM peak; `cpu_baseline` = numpy restatement of basic_python timed
on this box's host (1 core), plus the C/OpenMP oracle on all cores as `cpu_oracle`.

--impl reference: runs the UNMODIFIED reference (optimized/main.cu rebuilt for sm_100a as
oracle/_ref/cuIsing_ref) through its own CLI on the same lattice; under torchrun only rank 0
runs it (cuIsing_ref -d N drives all N GPUs from one process, as the reference does).  If the
binary is absent it times the CPU oracle port on all host cores instead and says so.
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


def default_rows(n):
    return 32768 if n >= 8 else 65536


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
    n = 4096
    sweeps = 24
    fl, mag, secs = numpy_basic.run(n=n, m=n, alpha=TEMP / numpy_basic.TCRIT, sweeps=sweeps, seed=1234)
    base = {"value": round(fl, 5), "unit": "flips/ns", "cores": 1, "host_cores": os.cpu_count(), "kind": "port",
            "sample": "numpy restatement of basic_python/ising_basic.py update_lattice, %dx%d, T=%.1f, %d sweeps, "
                      "%.1f s" % (n, n, TEMP, sweeps, secs)}
    L = ising_oracle.OracleLattice(X_COLS, 4096)
    L.init()
    t0 = time.perf_counter()
    L.sweeps(2, TEMP)
    dt = time.perf_counter() - t0
    L.close()
    orc = {"value": round(X_COLS * 4096 * 2 / dt * 1e-9, 4), "unit": "flips/ns", "cores": ising_oracle.num_threads(),
           "kind": "port", "sample": "bit-exact C/OpenMP oracle of optimized/, 4096 rows x 65536, 2 sweeps, %.1f s" % dt}
    return base, orc


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
        numa = bind_process_near_gpu(local)     # host staging buffers next to the GPU's PCIe root
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rows = a.rows_per_gpu or default_rows(world)
    S = a.sweeps_per_step
    eng = SlabEngine(X_COLS, rows, TEMP, a.seed, rank=rank, world=world, device=local, overlap=not a.no_overlap,
                     halo=a.halo)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    eng.init_lattice()
    for _ in range(a.warmup):
        eng.sweeps(S)
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = eng.eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        eng.sweeps(S)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.eng.launches - launches0
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    up, dn = eng.count_spins()
    flips_per_rank = float(X_COLS) * rows * S * a.steps
    value = flips_per_rank * world / (ms * 1e6)

    # ---- end to end through the C ABI with host buffers ------------------------------------
    e2e = None
    if not a.no_e2e:
        pitch = X_COLS // 64
        plane_bytes = rows * pitch * 4
        ksteps = max(1, min(a.steps, a.e2e_steps)) if a.e2e_steps else a.steps

        def pinned():
            t = torch.empty((rows, pitch), dtype=torch.int32).pin_memory()
            return t, t.numpy().view(np.uint32)
        b0, w0 = eng.eng.export_bits()

        def plain_e2e(nb, nw):
            """One step after the other, nothing overlapped (the transport-agnostic path; also the fallback)."""
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

        if world == 1:
            # replica batch: every step is one job of the batch = upload both planes of a lattice from
            # pinned memory, S sweeps, spin counts, download both planes; jobs alternate between two
            # contexts so the copies of one step overlap the kernels of the next (ising_run_batch_from_host)
            from ising_gpu_b200.engine import IsingEngine
            eng2 = IsingEngine(X_COLS, rows, seed=a.seed, temp=TEMP, first_device=local)
            keep, ins, outs = [], [], []
            for j in range(2):
                (tb, nb), (tw, nw) = pinned(), pinned()
                nb[:], nw[:] = b0, w0
                (tob, nob), (tow, now) = pinned(), pinned()
                keep += [tb, tw, tob, tow]
                ins.append((nb, nw)); outs.append((nob, now))
            del b0, w0

            def batch(k, download):
                return eng.eng.run_batch_from_host(eng2, [ins[j & 1][0] for j in range(k)],
                                                   [ins[j & 1][1] for j in range(k)], S,
                                                   [outs[j & 1][0] for j in range(k)] if download else None,
                                                   [outs[j & 1][1] for j in range(k)] if download else None,
                                                   first_it=eng.it + 1)

            def timed(download):
                batch(2, download)
                best = None
                for _ in range(2):                   # best of 2 runs of K steps (host-side transients)
                    barrier()
                    t0 = time.perf_counter()
                    counts = batch(ksteps, download)
                    barrier()
                    d = time.perf_counter() - t0
                    best = d if best is None else min(best, d)
                return best, counts
            dt_dl, counts = timed(True)              # D2H of the final planes as well (a checkpoint per step)
            assert len(set(counts)) == 1 and counts[0][0] + counts[0][1] == X_COLS * rows, counts
            assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
            dt, counts2 = timed(False)               # D2H of the step's observable only
            assert counts2 == counts
            eng2.close()
            api = ("ising_run_batch_from_host (C ABI, pinned host buffers): per step H2D of both planes, %d sweeps, "
                   "count kernels, D2H of the spin counts; two contexts, copies overlap the other step's kernels" % S)
        elif a.halo == "fused":
            # the same replica batch on every rank's slab (ising_run_slab_batch_from_host): per step H2D of the
            # slab's planes and white ghost rows, S sweeps with the fused halo push between the ranks, counts,
            # D2H of the planes; two contexts per rank so that copies overlap the other step's kernels
            with torch.cuda.stream(torch.cuda.Stream()):   # SlabEngine works on the stream current at construction
                eng2 = SlabEngine(X_COLS, rows, TEMP, a.seed, rank=rank, world=world, device=local, halo="fused")
            eng2.init_lattice()
            # one unpipelined step first: gives the ghost rows that belong to the uploaded state and the
            # result the pipelined steps must reproduce
            it0 = eng.it
            eng.eng.import_bits(b0, w0)
            eng.refresh_ghosts()
            ghosts = eng.white_ghost_rows()
            eng.sweeps(S)
            want_b, want_w = eng.eng.export_bits()
            want_counts = eng.count_spins()
            eng.it = eng2.it = it0
            keep, ins, outs, ready = [], [], [], True
            try:        # pinned host memory is the one local resource that can run out here
                for j in range(2):
                    (tb, nb), (tw, nw) = pinned(), pinned()
                    nb[:], nw[:] = b0, w0
                    (tob, nob), (tow, now) = pinned(), pinned()
                    keep += [tb, tw, tob, tow]
                    ins.append((nb, nw)); outs.append((nob, now))
                tg = torch.empty((2, pitch), dtype=torch.int32).pin_memory()
                ng = tg.numpy().view(np.uint32)
                ng[:] = ghosts
            except RuntimeError:
                ready = False
            t = torch.tensor([float(ready)], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)        # every rank enters the batch, or none does
            ready = bool(t.item())
            dt, ok = 1.0, False
            if ready:
                del b0, w0

                def batch(k, download):
                    return eng.run_batch_from_host(eng2, [ins[j & 1][0] for j in range(k)],
                                                   [ins[j & 1][1] for j in range(k)], [ng] * k, S,
                                                   [outs[j & 1][0] for j in range(k)] if download else None,
                                                   [outs[j & 1][1] for j in range(k)] if download else None)

                def timed(download):
                    batch(2, download)
                    best = None
                    for _ in range(2):               # best of 2 runs of K steps (host-side transients); the ranks
                        barrier()                    # agree on the maximum of their best times below
                        t0 = time.perf_counter()
                        counts = batch(ksteps, download)
                        barrier()
                        d = time.perf_counter() - t0
                        best = d if best is None else min(best, d)
                    return best, counts
                dt_dl, counts = timed(True)          # D2H of the final planes as well (a checkpoint per step)
                ok = all(c == want_counts for c in counts) and all(
                    np.array_equal(ob, want_b) and np.array_equal(ow, want_w) for ob, ow in outs[:min(2, ksteps)])
                dt, counts2 = timed(False)           # D2H of the step's observable only
                ok = ok and counts2 == counts
            del want_b, want_w
            t = torch.tensor([dt, -float(ok), dt_dl if ready else 1.0], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, ok, dt_dl = float(t[0].item()), float(t[1].item()) == -1.0, float(t[2].item())   # ok on every rank
            eng2.close()
            api = ("ising_run_slab_batch_from_host (C ABI, pinned host buffers): per step and rank H2D of the slab's planes "
                   "and white ghost rows, %d sweeps with the fused halo push, count kernels, D2H of the spin counts; "
                   "two contexts per rank, copies overlap the other step's kernels; result checked against the "
                   "unpipelined path" % S)
            if not ok:   # never report a number for a wrong result: measure the plain path instead and say so
                dt_dl = None
                if not ready:
                    keep, ins = [], []
                    (tb, nb), (tw, nw) = pinned(), pinned()
                    nb[:], nw[:] = b0, w0
                    ins.append((nb, nw))
                dt = plain_e2e(ins[0][0], ins[0][1])
                t = torch.tensor([dt], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
                api = (("PIPELINED RESULT MISMATCH" if ready else "pinned host memory exhausted") +
                       ", fell back to ising_import_bits + ising_sweeps + ising_export_bits + "
                       "ising_count_spins per step, not pipelined")
        else:
            (hb, nb), (hw, nw) = pinned(), pinned()
            nb[:], nw[:] = b0, w0
            del b0, w0
            dt = plain_e2e(nb, nw)
            dt_dl = None
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
            api = ("ising_import_bits + ising_sweeps + ising_export_bits + ising_count_spins per step "
                   "(C ABI, pinned host buffers), not pipelined")
        pipelined = dt_dl is not None
        h2d = (2 * plane_bytes + (2 * pitch * 4 if world > 1 and pipelined else 0)) * world
        e2e = {"value": round(float(X_COLS) * rows * S * ksteps * world / (dt * 1e9), 2), "unit": "flips/ns",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": (8 if pipelined else 2 * plane_bytes + 8) * world,
               "steps": ksteps, "timer": "host monotonic clock around the synchronised region" +
               (", best of 2 runs of %d steps" % ksteps if pipelined else ""), "api": api,
               "numa_node_of_rank0": numa}
        if pipelined:   # the same batch with the final lattice of every step downloaded as well
            e2e["with_lattice_download"] = {"value": round(float(X_COLS) * rows * S * ksteps * world / (dt_dl * 1e9), 2),
                                            "unit": "flips/ns", "h2d_bytes_per_step": h2d,
                                            "d2h_bytes_per_step": (2 * plane_bytes + 8) * world}

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
                "note": "HBM-nominal path, but the bit-exact Philox4x32-10 per site makes it bound by the "
                        "FMA-heavy pipe (IMAD.WIDE, 4 cycles/warp-instr): see DESIGN.md and profiles/"}
        out = {"metric": "spin-flip attempts/ns", "value": round(value, 2), "unit": "flips/ns", "n_gpus": world,
               "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms / a.steps, 4), "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
               "config": {"workload": "checkerboard Metropolis, %dx%d lattice (%d rows per GPU), T=%.1f, seed %d, "
                                      "%d sweeps per step" % (rows * world, X_COLS, rows, TEMP, a.seed, S),
                          "lattice_bytes_per_gpu": rows * X_COLS // 8, "l2_flush": "inputs larger than L2 "
                          "(%d MiB of spins per GPU vs 126 MB L2)" % (rows * X_COLS // 8 >> 20),
                          "halo": "none" if world == 1 else (
                              "fused: update kernel stores its edge rows into the neighbours' ghost rows (CUDA IPC "
                              "peer memory over NVLink) and signals device-side counters" if a.halo == "fused" else
                              "NCCL send/recv of 1 row per direction per colour, " +
                              ("overlapped with the interior update" if not a.no_overlap else "serialised"))},
               "roofline": roof, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
               "final_counts": {"up": up if world == 1 else None, "down": dn if world == 1 else None}}
        if world == 1 and not a.no_cpu:
            out["cpu_baseline"], out["cpu_oracle"] = cpu_baselines()
    if world > 1:
        # total spin counts as a cross-rank checksum
        t = torch.tensor([up, dn], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        if rank == 0:
            out["final_counts"] = {"up": int(t[0]), "down": int(t[1])}
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    if rank == 0:
        _GUARD.emit(json.dumps(out))


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
            "config": {"workload": "checkerboard Metropolis, %dx%d lattice (%d rows per GPU), T=%.1f, seed %d, "
                                   "%d sweeps per step" % (rows * n, X_COLS, rows, TEMP, a.seed, S)}}
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
                call(max(1, a.warmup))
            ms, fl, fin = call(S * a.steps)
            base.update({"value": fl, "ms_per_step": round(ms / a.steps, 4), "dtype": "u64 (4 bit/spin)",
                         "reference_kind": "unmodified optimized/main.cu + utils.c, nvcc -O3 --use_fast_math for "
                                           "sm_100a, run through its own CLI (cuIsing_ref -d %d)" % n,
                         "final_counts": {"up": fin[0], "down": fin[1]} if fin else None,
                         "cpu_baseline": {"value": fl, "unit": "flips/ns", "cores": 0, "kind": "reference",
                                          "sample": "GPU reference binary, not a CPU run: %d sweeps of the full "
                                                    "workload" % (S * a.steps)},
                         "e2e": {"value": fl, "unit": "flips/ns", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
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

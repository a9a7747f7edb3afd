#!/usr/bin/env python
"""bench.py -- particle RK4 steps/s of the hot path on B200 (BASELINE.json metric).

Workload (config C5 of BASELINE.json): pathline, 64 M seeds (strong scaling:
64 M / N per GPU), synthetic icosahedral-Voronoi MPAS mesh with 2,621,442 cells x 80 layers,
solid-body-rotation snapshots, RK4, dt = 120 s, depth 800 m.  One bench "step" = one
MOPS_RunPathLine-equivalent call (mops_pathline through the C ABI) over one snapshot interval
of `--interval-steps` RK4 steps (default 120 of the 720 a 1-day interval has, so that the
default run ends within minutes; throughput is per particle-step), chained the way the
reference's tutorial chains intervals (end points -> next seeds, re-located), while the NEXT
snapshot is uploaded + preprocessed on the side stream from pinned host memory
(double-buffered H2D).  Executed particle-steps (alive at step start) are counted by the
kernel itself.

  value        : device-resident arm (particles/outputs stay in HBM), K timed steps, max over ranks
  e2e          : same call with HOST buffers (pinned): seeds/depths H2D and recorded
                 trajectories D2H inside the timed region
  roofline     : k_advect launches only (CUDA events on the launching stream, from the C ABI's
                 stats), algorithmic bytes per step of SURVEY.md 8(d)
  cpu_baseline : the reference's own TBB/CPU implementation (oracle/_ref, compiled unmodified)
                 on the host cores, bounded sample (N = 1, rank 0 only)

`--impl reference` times only the reference CPU arm and prints the same JSON shape.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT = 120
DEPTH = 800.0
METRIC = "particle RK4 steps/sec (pathline, executed particle-steps)"
UNIT = "particle-steps/s"


def algorithmic_bytes_per_step(L: int, pathline: bool) -> int:
    """int32 indices, fp64 payload, nv = 6, minimal K = 2*ceil(log2(L-1)) + 2."""
    K = 2 * math.ceil(math.log2(L - 1)) + 2
    per_eval = (172 + 2 * (48 * K + 384)) if pathline else (556 + 48 * K)
    return 4 * per_eval + 196 + 64


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def make_snapshot_host(mesh, L, speed, tilt, pinned_alloc):
    """zonal / meridional / layerThickness / bottomDepth of a solid-body snapshot into (pinned) host arrays."""
    from mops_b200 import synthetic as S
    snap = S.solid_body_snapshot(mesh, L, speed, tilt=tilt)
    bufs = {}
    for k, a in (("zonal", snap.zonal), ("merid", snap.meridional), ("thick", snap.layer_thickness), ("bottom", snap.bottom_depth)):
        b = pinned_alloc(a.shape)
        b[...] = a
        bufs[k] = b
    del snap
    return bufs


def run_cpu_reference(level, L, n_particles, interval_steps, threads=None, target_seconds=12.0, repeats=1, log=None):
    """Times the reference's TBB/CPU PathLine on the host cores (oracle/_ref).  Returns dict or None."""
    from mops_b200 import synthetic as S
    try:
        from oracle import ref_oracle as R
        have_ref = R.available()
    except Exception:
        have_ref = False
    mesh = S.icosahedral_mesh(level)
    s0 = S.solid_body_snapshot(mesh, L, 0.02, tilt=0.3)
    s1 = S.solid_body_snapshot(mesh, L, 0.025, tilt=0.31)
    seeds_all = S.uniform_sphere_seeds(n_particles, 20261018 + 5)
    duration = DT * interval_steps
    if have_ref:
        cores = R.max_threads() if threads is None else threads
        R.set_threads(cores)
        o = R.RefOracle(mesh, [s0, s1])
        o.activate(0, 1)
        # probe to size the sample for ~target_seconds
        probe_n = min(n_particles, 20000)
        r = o.pathline(seeds_all[:probe_n], DT, duration, min(3600, duration), depth=DEPTH)
        rate = probe_n * interval_steps / max(r["seconds"], 1e-6)
        n = int(min(n_particles, max(probe_n, rate * target_seconds / interval_steps)))
        times = []
        for _ in range(repeats):
            r = o.pathline(seeds_all[:n], DT, duration, min(3600, duration), depth=DEPTH)
            times.append(r["seconds"])
        o.close()
        kind = "reference"
    else:
        from oracle import port_oracle as P
        cores = 1
        p0, p1 = P.prepare(mesh, s0), P.prepare(mesh, s1)
        n = min(n_particles, 20000)
        cell0 = P.locate(mesh, seeds_all[:n])
        times = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            P.pathline(mesh, p0, p1, seeds_all[:n], cell0, DT, duration, duration, depth=DEPTH, log_cells=False)
            times.append(time.perf_counter() - t0)
        kind = "port"
    # executed particle-steps of the sample: counted with the scalar port (bit-identical decisions)
    from oracle import port_oracle as P
    p0, p1 = P.prepare(mesh, s0), P.prepare(mesh, s1)
    sub = min(n, 20000)
    cell0 = P.locate(mesh, seeds_all[:sub])
    pr = P.pathline(mesh, p0, p1, seeds_all[:sub], cell0, DT, duration, duration, depth=DEPTH, log_cells=False)
    alive_frac = float(pr["steps_alive"].sum()) / float(sub * interval_steps)
    steps = n * interval_steps * alive_frac
    return {"kind": kind, "cores": int(cores), "times": times, "particle_steps": steps, "n": n,
            "sample": f"pathline, {n} seeds x {interval_steps} RK4 steps, {mesh.n_cells}-cell x {L}-layer mesh "
                      f"(same generator, level {level}), dt={DT}s, depth {DEPTH:.0f} m; includes the reference's host KD-tree "
                      f"lookup and line assembly (MOPS_RunPathLine wall time); executed fraction {alive_frac:.4f}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--level", type=int, default=9, help="icosahedral bisection level (9 = 2,621,442 cells)")
    ap.add_argument("--layers", type=int, default=80)
    ap.add_argument("--sweep-host-chunk", type=str, default="", help="experiment: comma list of HOST-mode chunk sizes to time")
    ap.add_argument("--particles", type=int, default=64_000_000, help="TOTAL seeds over all GPUs (strong scaling)")
    ap.add_argument("--interval-steps", type=int, default=120, help="RK4 steps per snapshot interval (720 = 1 day)")
    ap.add_argument("--cpu-level", type=int, default=7)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sort", action="store_true")
    ap.add_argument("--snapshot-allgather", action="store_true",
                    help="N > 1: every rank uploads 1/N of the next snapshot over PCIe and an NCCL all-gather over NVLink "
                         "completes it on every GPU (default: every rank uploads the whole snapshot itself)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W, K = max(args.warmup, 0), max(args.steps, 1)
    L = args.layers
    workload = (f"C5 pathline: {args.particles} seeds total, icosahedral level {args.level} "
                f"({10 * 4 ** args.level + 2} cells) x {L} layers, RK4 dt={DT}s, depth {DEPTH:.0f} m, "
                f"{args.interval_steps} RK4 steps per snapshot interval, next snapshot double-buffered H2D")

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        res = run_cpu_reference(args.cpu_level, L, 2_000_000, args.interval_steps, target_seconds=10.0, repeats=W + K)
        t = res["times"][W:]
        ms = 1e3 * float(np.mean(t))
        val = res["particle_steps"] / (ms / 1e3)
        line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "impl": "reference",
                "config": {"workload": workload, "reference_sample": res["sample"]},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    from mops_b200 import capi, synthetic as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    t_setup = time.perf_counter()
    mesh = S.icosahedral_mesh(args.level)
    eng = capi.Engine(local_rank)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.set_mesh(mesh)

    def pinned(shape):
        return torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()

    # distinct host snapshots (pinned); snapshot s of the chain re-uses ring[s % len(ring)].  Two when host
    # memory allows (all ranks of the box pin theirs at once), else one.
    snap_host_bytes = 3 * mesh.n_cells * L * 8
    ring_n = 2
    try:
        avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
        if world * (2 * snap_host_bytes + 8e9) > 0.6 * avail:
            ring_n = 1
    except Exception:
        pass
    ring = [make_snapshot_host(mesh, L, 0.02 * (1 + 0.5 * math.sin(2 * math.pi * s / 30)), 0.3 + 0.01 * s, pinned)
            for s in range(ring_n)]

    # N > 1, opt-in: partitioned upload + all-gather (SURVEY 8e "broadcast of each new snapshot").  Each rank keeps only its
    # 1/N chunk of the concatenated cell-major fields in pinned memory; the chunk goes over this rank's PCIe link on a
    # private torch stream, an all-gather on a private NCCL communicator assembles the snapshot in HBM (double-buffered:
    # the engine's side stream copies out of it asynchronously), and the engine takes DEVICE pointers.
    use_ag = bool(args.snapshot_allgather) and world > 1
    if use_ag:
        from mops_b200 import sharding as _sh
        nfield = mesh.n_cells * L
        ag_total = 3 * nfield + mesh.n_cells
        ag_chunk, _, _ = _sh.snapshot_part_bounds(ag_total, rank, world)
        ring_parts = []
        for h in ring:
            part_h = pinned((ag_chunk,))
            _sh.pack_snapshot_part([h["zonal"], h["merid"], h["thick"], h["bottom"]], rank, world, out=part_h)
            ring_parts.append(torch.from_numpy(part_h))
        ring = None  # the whole-snapshot host copies are not needed in this mode
        ag_group = dist.new_group(backend="nccl")
        ag_stream = torch.cuda.Stream()
        ag_full = [torch.empty(ag_chunk * world, dtype=torch.float64, device=dev) for _ in range(2)]
        ag_part = torch.empty(ag_chunk, dtype=torch.float64, device=dev)
        ag_event = [torch.cuda.Event() for _ in range(2)]

    def upload(slot, s, async_):
        if use_ag:
            k = s % 2
            with torch.cuda.stream(ag_stream):
                ag_part.copy_(ring_parts[s % ring_n], non_blocking=True)
                dist.all_gather_into_tensor(ag_full[k], ag_part, group=ag_group)
                ag_event[k].record(ag_stream)
            eng.side_wait_event(ag_event[k].cuda_event)
            base = ag_full[k].data_ptr()
            eng.set_snapshot_raw(slot, L, base, base + 8 * nfield, base + 16 * nfield, base + 24 * nfield, None, async_=async_)
            return
        h = ring[s % ring_n]
        eng.set_snapshot_raw(slot, L, h["zonal"].ctypes.data, h["merid"].ctypes.data, h["thick"].ctypes.data,
                             h["bottom"].ctypes.data, None, async_=async_)

    upload(0, 0, False)
    upload(1, 1, False)

    # seeds: uniform on the sphere |lat| < 80 deg, rank r keeps its longitude sector
    n_total = args.particles
    seeds_all = S.uniform_sphere_seeds(n_total, 20261018 + 5)
    from mops_b200 import sharding
    seeds_np, _global_idx = sharding.shard_seeds(seeds_all, rank, world)
    del seeds_all
    n = seeds_np.shape[0]
    duration = DT * args.interval_steps
    record_t = min(3600, duration)  # hourly records
    each = duration // record_t

    xyz = torch.from_numpy(seeds_np).to(dev)
    depth = torch.full((n,), DEPTH, dtype=torch.float32, device=dev)
    out_pos = torch.empty((n, each, 3), dtype=torch.float64, device=dev)
    out_vel = torch.empty((n, each, 3), dtype=torch.float64, device=dev)
    cfg = capi.TrajCfg(capi.METHOD_RK4, capi.DIR_FORWARD, DT, duration, record_t, capi.MEM_DEVICE, 0 if args.no_sort else 1)
    io = capi.TrajIO(n, xyz.data_ptr(), depth.data_ptr(), None, out_pos.data_ptr(), out_vel.data_ptr(), None, None, None, None, None)
    counts = sharding.all_counts(n, world, device=dev)
    setup_s = time.perf_counter() - t_setup

    def one_step(i, io_, cfg_):
        # next snapshot: async H2D + device preprocessing on the side stream (double buffering)
        upload((i + 2) % 3, i + 2, True)
        st = eng.traj_device(True, (i % 3, (i + 1) % 3), cfg_, io_, want_stats=True)
        if world > 1:
            # the one exchange of the path: end points gathered to rank 0 over NCCL/NVLink
            sharding.gather_rows(xyz, counts, rank, world, dst=0)
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reset_particles():
        xyz.copy_(torch.from_numpy(seeds_np))
        depth.fill_(DEPTH)

    # ---- device-resident arm ---------------------------------------------------------------
    step_no = 0
    for _ in range(W):
        one_step(step_no, io, cfg)
        step_no += 1
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    psteps, kms = 0, []
    kern_steps = []
    launches0 = int(eng.info().total_launches)
    for _ in range(K):
        st = one_step(step_no, io, cfg)
        step_no += 1
        psteps += int(st.particle_steps)
        kms.append(float(st.kernel_ms))
        kern_steps.append(int(st.particle_steps))
    ev1.record()
    barrier()
    launches = int(eng.info().total_launches) - launches0  # our kernels only (CUB sort passes not counted)
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total, float(psteps)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, psteps_all = float(tmax[0]), float(tsum[1])
    else:
        psteps_all = float(psteps)
    value = psteps_all / (ms_total / 1e3)

    # ---- e2e arm: same call with HOST (pinned) buffers --------------------------------------
    e2e = None
    if not args.no_e2e:
        h_xyz = pinned((n, 3)); h_xyz[...] = seeds_np
        h_depth = torch.full((n,), DEPTH, dtype=torch.float32).pin_memory().numpy()
        h_pos = pinned((n, each, 3)); h_vel = pinned((n, each, 3))
        cfg_h = capi.TrajCfg(capi.METHOD_RK4, capi.DIR_FORWARD, DT, duration, record_t, capi.MEM_HOST, 0 if args.no_sort else 1)
        io_h = capi.TrajIO(n, h_xyz.ctypes.data, h_depth.ctypes.data, None, h_pos.ctypes.data, h_vel.ctypes.data,
                           None, None, None, None, None)
        h2d = n * 24 + n * 4
        d2h = n * 24 + n * 4 + 2 * n * each * 24

        def e2e_step(i):
            upload((i + 2) % 3, i + 2, True)
            return eng.traj_device(True, (i % 3, (i + 1) % 3), cfg_h, io_h, want_stats=True)

        for _ in range(min(W, 2)):
            e2e_step(step_no); step_no += 1
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ps = 0
        for _ in range(K):
            st = e2e_step(step_no); step_no += 1
            ps += int(st.particle_steps)
        e1.record()
        barrier()
        ms_e = e0.elapsed_time(e1)
        wall_e = (time.perf_counter() - t0) * 1e3
        ms_e = max(ms_e, wall_e)  # host-side staging between calls counts too
        te = torch.tensor([ms_e, float(ps)], dtype=torch.float64, device=dev)
        if world > 1:
            a = te.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
            b = te.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
            ms_e, ps_all = float(a[0]), float(b[1])
        else:
            ps_all = float(ps)
        e2e = {"value": ps_all / (ms_e / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": ms_e / K}

        if args.sweep_host_chunk:  # experiment: HOST-mode pipeline chunk size (particles; 0 = single pass)
            for c in [int(x) for x in args.sweep_host_chunk.split(",")]:
                os.environ["MOPS_HOST_CHUNK"] = str(c if c > 0 else 1 << 40)
                e2e_step(step_no); step_no += 1
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                kk = 0.0
                for _ in range(2):
                    st = e2e_step(step_no); step_no += 1
                    kk += st.kernel_ms
                torch.cuda.synchronize()
                if rank == 0:
                    print(f"[sweep] host_chunk={c} ms_per_step={(time.perf_counter() - t0) * 500:.1f} kernel_ms={kk / 2:.1f}",
                          file=sys.stderr, flush=True)
            os.environ.pop("MOPS_HOST_CHUNK", None)

    # ---- roofline of the dominant kernel (k_advect<6,true>) ----------------------------------
    peak, peak_src = load_peaks()
    bps = algorithmic_bytes_per_step(L, True)
    ach = [bps * s / (m / 1e3) / 1e9 for s, m in zip(kern_steps, kms) if m > 0]
    achieved = float(np.mean(ach)) if ach else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("k_advect_pathline_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k_advect<6,true>", "peak_source": peak_src,
                "algorithmic_bytes_per_particle_step": bps, "kernel_ms_per_launch": float(np.mean(kms)),
                "kernel_share_of_step": float(np.sum(kms)) / (ev0.elapsed_time(ev1) if world == 1 else ms_total),
                "note": "logical bytes ignore L1/L2 reuse between particles sharing a cell; frac > 1 means the kernel is "
                        "cache-/fp64-bound, not HBM-bound (profiles/README.md): dram_* is the real DRAM rate"}
    if traffic and world == 1 and args.level == 9 and args.particles == 64_000_000 and args.interval_steps == 120 and kms:
        # traffic.json is an ncu capture of exactly this launch shape (64 M particles x 120 steps, level-9 mesh)
        roofline["dram_achieved"] = float(traffic) / (float(np.mean(kms)) / 1e3) / 1e9
        roofline["dram_frac"] = roofline["dram_achieved"] / peak

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            res = run_cpu_reference(args.cpu_level, L, 2_000_000, args.interval_steps, target_seconds=12.0)
            cpu = {"value": res["particle_steps"] / float(np.mean(res["times"])), "unit": UNIT, "cores": res["cores"],
                   "kind": res["kind"], "sample": res["sample"]}
        except Exception as ex:  # the baseline is a reported number; never fail the bench on it
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": f"failed: {ex}"}

    if rank == 0:
        info = eng.info()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "particles_total": n_total, "particles_this_rank": n,
                           "cells": mesh.n_cells, "layers": L, "interval_steps": args.interval_steps,
                           "l2_policy": "inputs larger than L2 (2 x 16.8 GB snapshots + particles); no flush needed",
                           "semantics": "reference (RK4 stages in the start-of-step cell; particles stop at their first failed stage)",
                           "sorted_particles": not args.no_sort, "setup_seconds": setup_s,
                           "mesh_bytes": int(info.mesh_bytes), "snapshot_bytes": int(info.snapshot_bytes[0]),
                           "parallelism": f"particles sharded by longitude sector over {world} GPU(s), mesh+snapshots replicated",
                           "snapshot_distribution": ("1/N PCIe upload per rank + NCCL all-gather" if use_ag
                                                     else "whole snapshot uploaded by every rank")},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""bench.py -- particle RK4 steps/s of the hot path on B200 (BASELINE.json metric).

Workload (config C5 of BASELINE.json): pathline, 64 M seeds (strong scaling:
64 M / N per GPU), synthetic icosahedral-Voronoi MPAS mesh with 2,621,442 cells x 80 layers,
solid-body-rotation snapshots, RK4, dt = 120 s, depth 800 m.  One bench "step" = one
MOPS_RunPathLine-equivalent call (mops_pathline through the C ABI) over one snapshot interval
of `--interval-steps` RK4 steps (default 120 of the 720 a 1-day interval has, so that the
default run ends within minutes; throughput is per particle-step), while the NEXT snapshot is
uploaded + preprocessed on the side stream from pinned host memory (H2D double-buffered two intervals ahead).

Every step integrates the SAME fresh seed set (particles are reset at the start of the step, a
device-to-device copy inside the timed region): under the reference's semantics a particle stops
for good at its first failed RK4 stage, so a chain that fed end points back in would execute a
shrinking fraction of its steps and `value` would depend on --steps.  Executed particle-steps
(alive at step start) are counted by the kernel itself; `config.executed_fraction` reports them.

  value        : device-resident arm (particles/outputs stay in HBM), K timed steps, max over ranks
  e2e          : the HOST-memory form of the same call (mops_pathline_submit / mops_traj_wait, pinned
                 buffers): seeds/depths H2D and end points + recorded trajectories D2H inside the timed
                 region, two staging sets so that the copies of one step overlap the kernels of the next
  roofline     : k_advect launches only (CUDA events on the launching stream, from the C ABI's stats).
                 The kernel is bound by the fp64 pipe, not by HBM (sorted particles share cells, L1 hit
                 98 %): `frac` is the fp64-pipe fraction, the contract's HBM figures sit beside it
  cpu_baseline : the reference's own TBB/CPU implementation (oracle/_ref, compiled unmodified)
                 on the host cores, bounded sample (N = 1, rank 0 only); `parity_sample` compares the
                 GPU on exactly that sample with the reference

`--impl reference` times only the reference CPU arm and prints the same JSON shape.
"""
from __future__ import annotations

import argparse
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
R_SEED = 6371010.0


def layer_levels_contract(L: int) -> int:
    """SURVEY.md 8(d) contract figure: minimal K = 2*ceil(log2(L-1)) + 2 zTop levels per vertex"""
    return 2 * math.ceil(math.log2(L - 1)) + 2


def algorithmic_bytes_per_step(L: int, pathline: bool, K: int) -> int:
    """int32 indices, fp64 payload, nv = 6, K zTop levels read per vertex and evaluation"""
    per_eval = (172 + 2 * (48 * K + 384)) if pathline else (556 + 48 * K)
    return 4 * per_eval + 196 + 64


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(d.get("sm_max_mhz", 1965.0))
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)", 1965.0


def load_profile(name):
    p = os.path.join(ROOT, "profiles", name)
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


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


def bind_to_gpu_numa_node(device: int):
    """Pin this process to the CPU cores of the NUMA node its GPU hangs off, BEFORE any pinned host memory is allocated
    (first touch puts the pages there): with one process per GPU on a two-socket box the snapshot uploads and the trajectory
    copy-backs otherwise cross the socket interconnect for half of the ranks.  Best effort: returns the node or None."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(device), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip()
        bdf = out.lower()
        if bdf.startswith("0000") and len(bdf.split(":")[0]) == 8:   # nvidia-smi prints an 8-digit domain, sysfs a 4-digit one
            bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = set(os.sched_getaffinity(0)) & set(cpus)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


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


def cpu_sample_inputs(level, L):
    """mesh, the two snapshots and the seed pool of the CPU-baseline / parity sample (same generators as the GPU workload)"""
    from mops_b200 import synthetic as S
    mesh = S.icosahedral_mesh(level)
    s0 = S.solid_body_snapshot(mesh, L, 0.02, tilt=0.3)
    s1 = S.solid_body_snapshot(mesh, L, 0.025, tilt=0.31)
    return mesh, s0, s1


def run_cpu_reference(level, L, n_particles, interval_steps, threads=None, target_seconds=12.0, repeats=1, keep_lines=False):
    """Times the reference's TBB/CPU PathLine on the host cores (oracle/_ref).  Returns dict."""
    from mops_b200 import synthetic as S
    try:
        from oracle import ref_oracle as R
        have_ref = R.available()
    except Exception:
        have_ref = False
    mesh, s0, s1 = cpu_sample_inputs(level, L)
    seeds_all = S.uniform_sphere_seeds(n_particles, 20261018 + 5)
    duration = DT * interval_steps
    record_t = min(3600, duration)
    lines = None
    if have_ref:
        # all host cores, set explicitly: under torchrun OMP_NUM_THREADS=1 is exported and omp_get_max_threads() obeys it
        cores = int(threads or os.cpu_count() or 1)
        R.set_threads(cores)
        o = R.RefOracle(mesh, [s0, s1])
        o.activate(0, 1)
        probe_n = min(n_particles, 20000)
        r = o.pathline(seeds_all[:probe_n], DT, duration, record_t, depth=DEPTH)
        rate = probe_n * interval_steps / max(r["seconds"], 1e-6)
        n = int(min(n_particles, max(probe_n, rate * target_seconds / interval_steps)))
        times = []
        for _ in range(repeats):
            r = o.pathline(seeds_all[:n], DT, duration, record_t, depth=DEPTH)
            times.append(r["seconds"])
        if keep_lines:
            lines = {"points": r["points"], "velocity": r["velocity"]}
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
            P.pathline(mesh, p0, p1, seeds_all[:n], cell0, DT, duration, record_t, depth=DEPTH, log_cells=False)
            times.append(time.perf_counter() - t0)
        kind = "port"
    # executed particle-steps + per-step cell ids of a subsample: the scalar port (bit-identical to the reference, logs ids)
    from oracle import port_oracle as P
    p0, p1 = P.prepare(mesh, s0), P.prepare(mesh, s1)
    sub = min(n, 20000)
    cell0 = P.locate(mesh, seeds_all[:sub])
    pr = P.pathline(mesh, p0, p1, seeds_all[:sub], cell0, DT, duration, record_t, depth=DEPTH, log_cells=True)
    alive_frac = float(pr["steps_alive"].sum()) / float(sub * interval_steps)
    steps = n * interval_steps * alive_frac
    return {"kind": kind, "cores": int(cores), "times": times, "particle_steps": steps, "n": n, "sub": sub,
            "mesh": mesh, "snaps": (s0, s1), "seeds": seeds_all[:n], "lines": lines, "port": pr, "record_t": record_t,
            "duration": duration,
            "sample": f"pathline, {n} seeds x {interval_steps} RK4 steps, {mesh.n_cells}-cell x {L}-layer mesh "
                      f"(same generator, level {level}), dt={DT}s, depth {DEPTH:.0f} m; includes the reference's host KD-tree "
                      f"lookup and line assembly (MOPS_RunPathLine wall time); executed fraction {alive_frac:.4f}"}


def parity_on_sample(res, device):
    """GPU vs the reference on exactly the sample the CPU arm integrated (outside every timed region): recorded
    positions / velocities of all its seeds against the compiled reference's lines, per-step cell ids of a subsample
    against the restatement's log, near-edge count (1e-12 rad band) from the diagnostic instantiation."""
    from mops_b200 import capi
    mesh, (s0, s1), seeds = res["mesh"], res["snaps"], res["seeds"]
    eng = capi.Engine(device)
    try:
        eng.set_mesh(mesh)
        eng.set_snapshot(0, s0)
        eng.set_snapshot(1, s1)
        g = eng.pathline(0, 1, seeds, DT, res["duration"], res["record_t"], depth=DEPTH, want_attr=False)
        sub = res["sub"]
        gs = eng.pathline(0, 1, seeds[:sub], DT, res["duration"], res["record_t"], depth=DEPTH, want_attr=False, log_cells=True,
                          near_edge=True)
        out = {"n": int(seeds.shape[0]), "cell_id_subsample": int(sub),
               "cell_mismatch": int(np.count_nonzero(gs["cell_log"] != res["port"]["cell_log"])),
               "near_edge_particles_subsample": int(gs["stats"].near_edge_particles)}
        if res["lines"] is not None:
            lines = eng.finalize_lines(seeds, g["raw_pos"], g["raw_vel"], pathline_mode=True)
            dx = np.linalg.norm(lines["points"] - res["lines"]["points"], axis=2)
            rv = res["lines"]["velocity"]
            dv = np.linalg.norm(lines["velocity"] - rv, axis=2) / np.maximum(np.linalg.norm(rv, axis=2), 1e-300)
            dv = np.where(np.linalg.norm(rv, axis=2) > 0, dv, 0.0)
            out.update({"against": "oracle/_ref (compiled reference) lines + oracle port cell log",
                        "max_dx_m": float(dx.max()), "max_rel_dv": float(dv.max()),
                        "bit_identical": bool(np.array_equal(lines["points"], res["lines"]["points"])
                                              and np.array_equal(lines["velocity"], rv))})
        else:
            dx = np.linalg.norm(gs["raw_pos"] - res["port"]["raw_pos"], axis=2)
            out.update({"against": "oracle port (reference not compiled on this box)", "max_dx_m": float(dx.max()),
                        "bit_identical": bool(np.array_equal(gs["raw_pos"], res["port"]["raw_pos"]))})
        return out, eng
    except Exception:
        eng.close()
        raise


def remap_secondary(eng, mesh, torch):
    """BASELINE metric's second half: remap pixels/s on C2 (level-7 mesh x 60 layers, depth 800 m), device-resident
    kernel time and end to end with the image copied back to host memory"""
    from mops_b200 import capi, synthetic as S
    s = S.solid_body_snapshot(mesh, 60, 0.5, tilt=0.3)
    eng.set_snapshot(2, s)
    out = {}
    for (w, h) in ((360, 180), (3600, 1800)):
        img = torch.empty((h, w, 4), dtype=torch.float64, device="cuda")
        cfg = capi.RemapCfg(w, h, -90.0, 90.0, -180.0, 180.0, 800.0, capi.MEM_DEVICE)
        for _ in range(3):
            eng.remap_device(2, cfg, img)
        kms = float(np.median([eng.remap_device(2, cfg, img).kernel_ms for _ in range(10)]))
        t0 = time.perf_counter()
        for _ in range(3):
            r = eng.remap(2, w, h, depth=800.0, want_attr=False, want_cells=False)
        e2e_ms = (time.perf_counter() - t0) / 3 * 1e3
        out[f"{w}x{h}"] = {"pixels_per_s": w * h / (kms / 1e3), "kernel_ms": kms, "pixels_per_s_e2e_host_image": w * h / (e2e_ms / 1e3),
                           "nan_pixels": int(r["stats"].nan_pixels)}
    out["config"] = "C2: 163,842 cells x 60 layers, depth 800 m, lat/lon full range; kernel = locate + interpolate per pixel"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--level", type=int, default=9, help="icosahedral bisection level (9 = 2,621,442 cells)")
    ap.add_argument("--layers", type=int, default=80)
    ap.add_argument("--particles", type=int, default=64_000_000, help="TOTAL seeds over all GPUs (strong scaling)")
    ap.add_argument("--interval-steps", type=int, default=120, help="RK4 steps per snapshot interval (720 = 1 day)")
    ap.add_argument("--cpu-level", type=int, default=7)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sort", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the near-edge pass, the parity sample and the remap figures")
    ap.add_argument("--chain", action="store_true", help="feed end points back in instead of resetting the particles every step")
    ap.add_argument("--snapshot-upload", default="auto", choices=["auto", "replicated", "allgather"],
                    help="N > 1: 'allgather' = every rank uploads 1/N of the next snapshot over PCIe and an NCCL all-gather over "
                         "NVLink completes it on every GPU; 'replicated' = every rank uploads the whole snapshot over its own PCIe "
                         "link; auto = replicated up to 4 GPUs, allgather beyond (measured on the 8 x B200 box, profiles/README.md)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W, K = max(args.warmup, 0), max(args.steps, 1)
    L = args.layers
    workload = (f"C5 pathline: {args.particles} seeds total, icosahedral level {args.level} "
                f"({10 * 4 ** args.level + 2} cells) x {L} layers, RK4 dt={DT}s, depth {DEPTH:.0f} m, "
                f"{args.interval_steps} RK4 steps per snapshot interval (C5 as written: 720), next snapshot double-buffered H2D, "
                f"{'chained end points' if args.chain else 'fresh seeds every step'}")

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
    from mops_b200 import capi, sharding, synthetic as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    t_setup = time.perf_counter()
    mesh = S.icosahedral_mesh(args.level)
    eng = capi.Engine(local_rank)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.set_mesh(mesh)

    def pinned(shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype, pin_memory=True).numpy()

    # distinct host snapshots (pinned); snapshot s of the chain re-uses ring[s % len(ring)].  Two when host
    # memory allows (all ranks of the box pin theirs at once), else one.
    # auto: replicated uploads up to 4 GPUs, all-gather from 8 on.  Measured on the 8 x B200 box (profiles/README.md): at N = 8
    # eight whole-snapshot uploads per interval (40 GB) crowd the host side of the e2e arm's own copies (e2e 18.3 G particle-steps/s
    # replicated vs 26.5 G all-gather, device-resident 30.9 vs 28.3); at N = 2 replicated wins both (8.59 / 8.20 vs 8.34 / 7.95)
    use_ag = world > 1 and (args.snapshot_upload == "allgather" or (args.snapshot_upload == "auto" and world > 4))
    snap_host_bytes = 3 * mesh.n_cells * L * 8
    ring_n = 2
    try:
        avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
        per_rank = (2 * snap_host_bytes / (world if use_ag else 1)) + 40e9 / world
        if world * per_rank > 0.6 * avail:
            ring_n = 1
    except Exception:
        pass
    ring = [make_snapshot_host(mesh, L, 0.02 * (1 + 0.5 * math.sin(2 * math.pi * s / 30)), 0.3 + 0.01 * s, pinned)
            for s in range(ring_n)]

    # N > 1: partitioned upload + all-gather (SURVEY 8e "broadcast of each new snapshot").  Each rank keeps only its 1/N chunk
    # of the concatenated cell-major fields in pinned memory; the chunk goes over this rank's PCIe link on a private torch
    # stream, an all-gather on a private NCCL communicator assembles the snapshot in HBM (double-buffered: the engine's side
    # stream copies out of it asynchronously), and the engine takes DEVICE pointers.
    nfield = mesh.n_cells * L
    if use_ag:
        ag_total = 3 * nfield + mesh.n_cells
        ag_chunk, _, _ = sharding.snapshot_part_bounds(ag_total, rank, world)
        ring_parts = []
        for h in ring:
            part_h = pinned((ag_chunk,))
            sharding.pack_snapshot_part([h["zonal"], h["merid"], h["thick"], h["bottom"]], rank, world, out=part_h)
            ring_parts.append(torch.from_numpy(part_h))
        ring = None  # the whole-snapshot host copies are not needed in this mode
        ag_group = dist.new_group(backend="nccl")
        ag_stream = torch.cuda.Stream()
        ag_full = [torch.empty(ag_chunk * world, dtype=torch.float64, device=dev) for _ in range(2)]
        ag_part = torch.empty(ag_chunk, dtype=torch.float64, device=dev)
        ag_event = [torch.cuda.Event() for _ in range(2)]
        ag_slot = [None, None]  # engine slot that consumed ag_full[k] last

    def upload(slot, s, async_):
        if use_ag:
            k = s % 2
            if ag_slot[k] is not None:
                # the engine's side stream may still be copying out of ag_full[k] (the previous upload that used it):
                # its `ready` event covers those copies, so wait for it before the all-gather overwrites the buffer
                eng.snapshot_wait(ag_slot[k])
            with torch.cuda.stream(ag_stream):
                ag_part.copy_(ring_parts[s % ring_n], non_blocking=True)
                dist.all_gather_into_tensor(ag_full[k], ag_part, group=ag_group)
                ag_event[k].record(ag_stream)
            eng.side_wait_event(ag_event[k].cuda_event)
            base = ag_full[k].data_ptr()
            eng.set_snapshot_raw(slot, L, base, base + 8 * nfield, base + 16 * nfield, base + 24 * nfield, None, async_=async_)
            ag_slot[k] = slot
            return
        h = ring[s % ring_n]
        eng.set_snapshot_raw(slot, L, h["zonal"].ctypes.data, h["merid"].ctypes.data, h["thick"].ctypes.data,
                             h["bottom"].ctypes.data, None, async_=async_)

    # four resident slots: interval i integrates between slots i % 4 and (i + 1) % 4 while snapshot i + 3 is uploaded and
    # preprocessed on the side stream, i.e. every snapshot has TWO intervals to arrive (at N = 8 an upload takes about as long
    # as one 215 ms interval, and with one interval of slack the kernels waited ~20 ms per step for it)
    NS = 4
    upload(0, 0, False)
    upload(1, 1, False)
    upload(2, 2, False)

    # seeds: uniform on the sphere |lat| < 80 deg.  Sharding: the seed set is sorted along the mesh's
    # Morton curve (each seed's cell from the engine's own point location, cells ranked along the curve) and cut into
    # `world` equal contiguous blocks, so every rank gets the same count for ANY seed distribution and a spatially compact
    # working set; `perm` is the caller-order index of each local seed (results scatter back by it).
    n_total = args.particles
    seeds_all = S.uniform_sphere_seeds(n_total, 20261018 + 5)
    if world > 1:
        all_dev = torch.from_numpy(seeds_all).to(dev)
        key = eng.order_key(all_dev).to(torch.int64) & 0xFFFFFFFF  # unsigned: -1 (no cell) sorts last, as in the engine
        order = torch.argsort(key, stable=True)
        lo, hi = sharding.block_bounds(n_total, rank, world)
        perm = order[lo:hi].to(torch.int32).contiguous()             # caller index of each local seed
        seeds0 = all_dev[perm.long()].contiguous()
        del all_dev, key, order
        torch.cuda.empty_cache()
        # NCCL communicator inside the library (the id travels over torch.distributed); the per-interval gather of the
        # recorded trajectories + end points to rank 0, in caller order, is the product's mops_dist_gather_traj
        uid = [eng.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.dist_create(rank, world, uid[0])
        comm_stream = torch.cuda.Stream()
        eng.dist_set_stream(comm_stream.cuda_stream)
    else:
        perm = None
        seeds0 = torch.from_numpy(seeds_all).to(dev)
    del seeds_all
    n = int(seeds0.shape[0])
    duration = DT * args.interval_steps
    record_t = min(3600, duration)  # hourly records
    each = duration // record_t

    xyz = seeds0.clone()
    depth = torch.full((n,), DEPTH, dtype=torch.float32, device=dev)
    sort = 0 if args.no_sort else 1
    cfg = capi.TrajCfg(capi.METHOD_RK4, capi.DIR_FORWARD, DT, duration, record_t, capi.MEM_DEVICE, sort)
    # two output sets when N > 1: the gather of interval i (comm stream) overlaps the kernels of interval i+1
    n_sets = 2 if world > 1 else 1
    outs = [(torch.empty((n, each, 3), dtype=torch.float64, device=dev), torch.empty((n, each, 3), dtype=torch.float64, device=dev))
            for _ in range(n_sets)]
    ios = [capi.TrajIO(n, xyz.data_ptr(), depth.data_ptr(), None, o[0].data_ptr(), o[1].data_ptr(), None, None, None, None, None) for o in outs]
    out_pos, out_vel = outs[0]
    io = ios[0]
    counts = sharding.all_counts(n, world, device=dev)
    # one owner for the recorded trajectories needs 2 x n_total x each x 24 B on rank 0, and as much again for the receive
    # staging: gathered while that fits beside the snapshots (C5 at 120-step intervals: 25 + 21 GB), otherwise the records
    # stay sharded and only the end points travel (C5 at 720 steps would need 2 x 74 GB on one GPU)
    gather_records = world > 1 and (4 * n_total * each * 24) < 70e9
    if world > 1:
        ends = [(torch.empty((n, 3), dtype=torch.float64, device=dev), torch.empty((n,), dtype=torch.float32, device=dev)) for _ in range(2)]
        gathered = (([torch.empty((n_total, each, 3), dtype=torch.float64, device=dev) for _ in range(2)] if gather_records else [None, None])
                    + [torch.empty((n_total, 3), dtype=torch.float64, device=dev), torch.empty((n_total,), dtype=torch.float32, device=dev)]
                    if rank == 0 else [None] * 4)
        k_done = [torch.cuda.Event() for _ in range(2)]
        g_done = [torch.cuda.Event() for _ in range(2)]
        g_used = [False, False]
    setup_s = time.perf_counter() - t_setup

    def reset_particles():
        if not args.chain:
            xyz.copy_(seeds0)
            depth.fill_(DEPTH)

    def one_step(i, io_, cfg_):
        k = i % n_sets
        if world > 1 and g_used[k]:
            torch.cuda.current_stream().wait_event(g_done[k])  # output set k is free again once its gather has run
        reset_particles()
        # next snapshot: async H2D + device preprocessing on the side stream (double buffering)
        upload((i + 3) % NS, i + 3, True)
        st = eng.traj_device(True, (i % NS, (i + 1) % NS), cfg_, ios[k] if io_ is io else io_, want_stats=True)
        if world > 1:
            # the path's one exchange: recorded trajectories + end points to rank 0 in caller order, NCCL over NVLink
            # (variable-size send/recv inside the library), on the comm stream so that it overlaps the next interval
            ends[k][0].copy_(xyz); ends[k][1].copy_(depth)
            k_done[k].record()
            comm_stream.wait_event(k_done[k])
            eng.dist_gather_traj(0, counts, perm, each, pos=outs[k][0] if gather_records else None, vel=outs[k][1] if gather_records else None,
                                 xyz=ends[k][0], depth=ends[k][1], n_total=n_total,
                                 out_pos=gathered[0], out_vel=gathered[1], out_xyz=gathered[2], out_depth=gathered[3])
            g_done[k].record(comm_stream)
            g_used[k] = True
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    launches = int(eng.info().total_launches) - launches0  # our kernels only (CUB sort / select passes not counted)
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    ms_local = ms_total
    t = torch.tensor([ms_total, float(psteps)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, psteps_all = float(tmax[0]), float(tsum[1])
    else:
        psteps_all = float(psteps)
    value = psteps_all / (ms_total / 1e3)
    executed_fraction = psteps_all / (float(n_total) * args.interval_steps * K)
    per_step = torch.tensor(kern_steps, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(per_step, op=dist.ReduceOp.SUM)
    executed_fraction_per_step = [float(x) / (float(n_total) * args.interval_steps) for x in per_step.tolist()]

    # ---- e2e arm: the HOST-memory form (submit / wait, pinned buffers, two sets) ----------------
    e2e = None
    if not args.no_e2e:
        h_seeds = pinned((n, 3)); h_seeds[...] = seeds0.cpu().numpy()
        h_seeds_t = torch.from_numpy(h_seeds)
        cfg_h = capi.TrajCfg(capi.METHOD_RK4, capi.DIR_FORWARD, DT, duration, record_t, capi.MEM_HOST, sort)
        sets = []
        for _ in range(2):
            hx = pinned((n, 3)); hd = pinned((n,), torch.float32); hp = pinned((n, each, 3)); hv = pinned((n, each, 3))
            sets.append({"xyz": torch.from_numpy(hx), "depth": torch.from_numpy(hd), "keep": (hx, hd, hp, hv),
                         "io": capi.TrajIO(n, hx.ctypes.data, hd.ctypes.data, None, hp.ctypes.data, hv.ctypes.data,
                                           None, None, None, None, None)})
        h2d = n * 24 + n * 4
        d2h = n * 24 + n * 4 + 2 * n * each * 24

        def e2e_run(count):
            """`count` steps, software-pipelined over the two host/staging sets; returns executed particle-steps"""
            nonlocal step_no
            tickets, ps = [], 0
            for i in range(count):
                k = i & 1
                if i >= 2:
                    ps += int(eng.traj_wait(tickets[i - 2], 1).particle_steps)  # set k is free again
                hs = sets[k]
                if not args.chain or i < 2:
                    hs["xyz"].copy_(h_seeds_t); hs["depth"].fill_(DEPTH)        # this step's inputs, in host memory
                else:
                    eng.traj_wait(tickets[i - 1], 0)                            # chain: the previous step's end points
                    hs["xyz"].copy_(sets[k ^ 1]["xyz"]); hs["depth"].copy_(sets[k ^ 1]["depth"])
                upload((step_no + 3) % NS, step_no + 3, True)
                tickets.append(eng.traj_submit(True, (step_no % NS, (step_no + 1) % NS), cfg_h, hs["io"]))
                step_no += 1
            for tk in tickets[max(0, count - 2):]:
                ps += int(eng.traj_wait(tk, 1).particle_steps)
            return ps

        e2e_run(min(W, 2))
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ps = e2e_run(K)
        e1.record()
        barrier()
        wall_e = (time.perf_counter() - t0) * 1e3
        ms_e = max(e0.elapsed_time(e1), wall_e)  # the copies run on the engine's own streams: the host-side wait is the clock
        te = torch.tensor([ms_e, float(ps)], dtype=torch.float64, device=dev)
        if world > 1:
            a = te.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
            b = te.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
            ms_e, ps_all = float(a[0]), float(b[1])
        else:
            ps_all = float(ps)
        e2e = {"value": ps_all / (ms_e / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": ms_e / K,
               "api": "mops_pathline_submit / mops_traj_wait (MOPS_MEM_HOST, pinned buffers, two staging sets)"}
        del sets

    # ---- near-edge particles of this workload (diagnostic instantiation, outside the timed regions) -----------------
    near_edge = None
    if not args.no_secondary:
        try:  # a reported diagnostic: never lose the bench line over it
            xyz.copy_(seeds0); depth.fill_(DEPTH)
            cfg_e = capi.TrajCfg(capi.METHOD_RK4, capi.DIR_FORWARD, DT, duration, record_t, capi.MEM_DEVICE, sort, 1)
            torch.cuda.synchronize()
            st_e = eng.traj_device(True, (step_no % NS, (step_no + 1) % NS), cfg_e, io, want_stats=True)
            ne = float(st_e.near_edge_particles)
        except Exception as ex:
            print(f"[bench] near-edge pass failed on rank {rank}: {ex}", file=sys.stderr)
            ne = float("nan")
        te = torch.tensor([ne], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.SUM)
        near_edge = None if math.isnan(float(te[0])) else int(te[0])

    # ---- roofline of the dominant kernel ------------------------------------------------------------------------
    peak, peak_src, sm_max_mhz = load_peaks()
    info = eng.info()
    Kc = layer_levels_contract(L)
    bps_contract = algorithmic_bytes_per_step(L, True, Kc)
    bps_read = algorithmic_bytes_per_step(L, True, 2)  # the layer hint makes the kernel read K = 2 levels per vertex
    kern_rate = [s / (m / 1e3) for s, m in zip(kern_steps, kms) if m > 0]
    rate = float(np.mean(kern_rate)) if kern_rate else 0.0
    prof = load_profile("k_advect_profile.json") or {}
    same_shape = (world == 1 and args.level == 9 and args.particles == 64_000_000 and args.interval_steps == 120)
    fp64_per_step = prof.get("fp64_thread_inst_per_particle_step")
    fp64_peak = info.sm_count * 64 * sm_max_mhz * 1e6  # fp64 thread-instructions/s: 64 lanes per SM and clock (ncu: dfma peak_sustained)
    traffic = prof.get("dram_bytes_per_launch") if same_shape else None
    roofline = {
        "bound": "fp64",
        "achieved": (fp64_per_step * rate / 1e12) if fp64_per_step else None,
        "peak": fp64_peak / 1e12,
        "unit": "Tinst/s (fp64-pipe thread instructions; DFMA, DMUL, DADD, DSETP each take one slot)",
        "frac": (fp64_per_step * rate / fp64_peak) if fp64_per_step else None,
        "traffic": traffic,
        "kernel": prof.get("kernel", "void mops::k_advect<6, 1, 3, 0, 0, 1, 1, 1>(AdvectParams)"),
        "fp64_thread_inst_per_particle_step": fp64_per_step,
        "fp64_source": prof.get("source"),
        "peak_source": f"{info.sm_count} SMs x 64 fp64 lanes x {sm_max_mhz:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json)",
        "ncu": {k: prof.get(k) for k in ("fp64_pipe_active_pct_ncu", "l1_data_pipe_pct", "l1_hit_pct", "l2_hit_pct", "warps_active_pct",
                                         "registers_per_thread", "dram_bytes_per_particle_step", "thread_inst_per_particle_step")},
        "second_roof": "L1 data pipe (LSU write-back, 128 B/clk/SM): l1_data_pipe_pct of peak in the same capture; the kernel sits "
                       "between the two, see DESIGN.md section 5",
        "kernel_ms_per_launch": float(np.mean(kms)) if kms else None,
        "kernel_particle_steps_per_s": rate,
        "kernel_share_of_step": (float(np.sum(kms)) / ms_local) if kms else None,
        "hbm": {
            "note": "the contract's HBM roof (SURVEY 8d): logical bytes ignore L1/L2 reuse between the ~24 particles sharing a "
                    "cell, so frac > 1 only says the kernel is not HBM-bound; dram_* is the real DRAM rate (ncu)",
            "peak": peak, "unit": "GB/s", "peak_source": peak_src,
            "algorithmic_bytes_per_particle_step_contract": bps_contract, "contract_levels_per_vertex": Kc,
            "achieved_contract": bps_contract * rate / 1e9, "frac_contract": bps_contract * rate / 1e9 / peak,
            "algorithmic_bytes_per_particle_step_as_read": bps_read, "levels_per_vertex_as_read": 2,
            "achieved_as_read": bps_read * rate / 1e9, "frac_as_read": bps_read * rate / 1e9 / peak,
        },
    }
    if traffic and kms:
        launches_per_step = prof.get("launches_per_step", 3)
        roofline["hbm"]["dram_achieved"] = float(traffic) * launches_per_step / (float(np.mean(kms)) / 1e3) / 1e9
        roofline["hbm"]["dram_frac"] = roofline["hbm"]["dram_achieved"] / peak

    config = {"workload": workload, "particles_total": n_total, "particles_this_rank": n,
              "cells": mesh.n_cells, "layers": L, "interval_steps": args.interval_steps,
              "executed_fraction": executed_fraction,
              "executed_fraction_per_step": [round(x, 4) for x in executed_fraction_per_step] if args.chain else None,
              "near_edge_particles": near_edge,
              "l2_policy": "inputs larger than L2 (2 x 16.8 GB snapshots + particles); no flush needed",
              "semantics": "reference (RK4 stages in the start-of-step cell; particles stop at their first failed stage)",
              "sorted_particles": not args.no_sort, "setup_seconds": setup_s,
              "mesh_bytes": int(info.mesh_bytes), "snapshot_bytes": int(info.snapshot_bytes[0]),
              "parallelism": f"seeds sorted along the mesh's Morton curve and cut into {world} equal contiguous block(s), "
                             "mesh+snapshots replicated",
              "numa_node_of_rank0": numa_node,
              "trajectory_gather": (None if world == 1 else
                                    "records + end points to rank 0 in caller order, NCCL send/recv inside the library (mops_dist_gather_traj), "
                                    "overlapped with the next interval" if gather_records else
                                    "end points only: the records of this configuration would not fit one GPU and stay sharded"),
              "snapshot_distribution": ("1/N PCIe upload per rank + NCCL all-gather" if use_ag
                                        else "whole snapshot uploaded by every rank")}
    torch.cuda.synchronize()
    eng.close()
    del xyz, depth, out_pos, out_vel, outs, seeds0
    torch.cuda.empty_cache()

    # ---- CPU baseline + parity on its sample + remap figures (rank 0, N = 1 only) -----------------------------------
    cpu = parity = remap = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            res = run_cpu_reference(args.cpu_level, L, 2_000_000, args.interval_steps, target_seconds=12.0, keep_lines=True)
            cpu = {"value": res["particle_steps"] / float(np.mean(res["times"])), "unit": UNIT, "cores": res["cores"],
                   "kind": res["kind"], "sample": res["sample"]}
            if not args.no_secondary:
                parity, eng2 = parity_on_sample(res, local_rank)
                try:
                    remap = remap_secondary(eng2, res["mesh"], torch)
                finally:
                    eng2.close()
        except Exception as ex:  # the baseline is a reported number; never fail the bench on it
            if cpu is None:
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": f"failed: {ex}"}
            else:
                parity = {"error": str(ex)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if parity is not None:
            line["parity_sample"] = parity
        if remap is not None:
            line["remap_pixels_per_s"] = remap
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

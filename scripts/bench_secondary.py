#!/usr/bin/env python
"""Secondary measurements of BASELINE.json configs C2 (remap) and C3 (1 M-seed streamline) on one B200,
with the reference's CPU path beside them.  Prints one JSON object per measurement (not the driver's
bench contract -- that is bench.py); results are copied into profiles/."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mops_b200 import capi, synthetic as S  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def remap_c2(eng, cpu=True):
    m = S.icosahedral_mesh(7)
    s0 = S.solid_body_snapshot(m, 60, 0.5, tilt=0.3)
    eng.set_mesh(m)
    eng.set_snapshot(0, s0)
    out = []
    for (w, h) in ((360, 180), (3600, 1800)):
        img = torch.empty((h, w, 4), dtype=torch.float64, device="cuda")
        cfg = capi.RemapCfg(w, h, -90.0, 90.0, -180.0, 180.0, 800.0, capi.MEM_DEVICE)
        for _ in range(3):
            eng.remap_device(0, cfg, img)
        ks = [eng.remap_device(0, cfg, img).kernel_ms for _ in range(10)]
        kms = float(np.median(ks))
        # end to end with host buffers (image D2H inside)
        t0 = time.perf_counter()
        for _ in range(3):
            r = eng.remap(0, w, h, depth=800.0, want_attr=False, want_cells=False)
        e2e_ms = (time.perf_counter() - t0) / 3 * 1e3
        px = w * h
        rec = {"config": f"C2 remap {w}x{h}, 163,842 cells x 60 layers, depth 800 m", "pixels_per_s_device": px / (kms / 1e3),
               "kernel_ms": kms, "pixels_per_s_e2e_host_image": px / (e2e_ms / 1e3), "e2e_ms": e2e_ms,
               "roofline_frac_592B_per_pixel": px * 592 / (kms / 1e3) / 1e9 / PEAK, "nan_pixels": int(r["stats"].nan_pixels)}
        out.append(rec)
    if cpu:
        try:
            from oracle import ref_oracle as R
            if R.available():
                o = R.RefOracle(m, [s0])
                rr = o.remap(360, 180, depth=800.0)
                rr2 = o.remap(1200, 600, depth=800.0)
                o.close()
                out.append({"config": "C2 remap CPU reference (oracle/_ref, incl. its serial host KD-tree loop)", "cores": R.max_threads(),
                            "pixels_per_s_360x180": 360 * 180 / rr["seconds"], "pixels_per_s_1200x600": 1200 * 600 / rr2["seconds"]})
        except Exception as ex:
            out.append({"config": "C2 CPU reference", "error": str(ex)})
    return out


def stream_c3(eng, days=7):
    m = S.icosahedral_mesh(8)
    s0 = S.solid_body_snapshot(m, 60, 0.02, tilt=0.3)
    eng.set_mesh(m)
    eng.set_snapshot(0, s0)
    n = 1_000_000
    seeds = S.gaussian_seeds(n, 20261018)
    dev = torch.device("cuda")
    duration, record_t = 86400 * days, 3600
    each = duration // record_t
    xyz0 = torch.from_numpy(seeds).to(dev)
    xyz = xyz0.clone()
    depth = torch.full((n,), 800.0, dtype=torch.float32, device=dev)
    out_pos = torch.empty((n, each, 3), dtype=torch.float64, device=dev)
    out_vel = torch.empty((n, each, 3), dtype=torch.float64, device=dev)
    cfg = capi.TrajCfg(capi.METHOD_RK4, capi.DIR_FORWARD, 120, duration, record_t, capi.MEM_DEVICE, 1)
    io = capi.TrajIO(n, xyz.data_ptr(), depth.data_ptr(), None, out_pos.data_ptr(), out_vel.data_ptr(), None, None, None, None, None)
    res = []
    for it in range(3):
        xyz.copy_(xyz0); depth.fill_(800.0)
        st = eng.traj_device(False, (0, 0), cfg, io, want_stats=True)
        res.append((st.kernel_ms, st.particle_steps, st.alive_at_end, st.total_ms))
    kms, steps, alive, tot = res[-1]
    return [{"config": f"C3 streamline, 1M Gaussian seeds, 655,362 cells x 60 layers, dt 120 s, {days} days ({duration // 120} steps), RK4",
             "particle_steps": int(steps), "alive_at_end": int(alive), "kernel_ms": kms, "call_ms": tot,
             "particle_steps_per_s_kernel": steps / (kms / 1e3), "particle_steps_per_s_call": steps / (tot / 1e3),
             "roofline_frac_5172B_per_step": steps * 5172 / (kms / 1e3) / 1e9 / PEAK}]


if __name__ == "__main__":
    eng = capi.Engine(0)
    for rec in remap_c2(eng) + stream_c3(eng):
        print(json.dumps(rec))
    eng.close()

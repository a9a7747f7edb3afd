#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the REFERENCE ITSELF (oracle/_ref/libmops_ref.so =
YosefQiu/MOPS TBB backend compiled unmodified by oracle/build_ref.sh) on small seeded cases.

Run here (where /root/reference exists):   python tests/golden/make_golden.py
The committed vectors pin oracle/mops_oracle.c (tests/test_oracle_golden.py) on machines
where the reference cannot be built (the GPU box).  Inputs are regenerated from
mops_b200.synthetic with the parameters stored in each file, so only outputs are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import cases  # noqa: E402
from oracle import ref_oracle as R  # noqa: E402

LEVEL, NLEV = 3, 8
DT, DUR = 120, 7200


def main():
    assert R.available(), "build oracle/_ref first (oracle/build_ref.sh)"
    for variant in ("plain", "rich"):
        m = cases.mesh(LEVEL)
        s0, s1 = cases.snapshots(LEVEL, NLEV, variant)
        o = R.RefOracle(m, [s0, s1])
        out = {"level": LEVEL, "n_levels": NLEV, "variant": variant, "dt": DT, "duration": DUR}
        for sid in (0, 1):
            p = o.prepared(sid)
            out[f"ztop_vertex_{sid}"] = p["ztop_vertex"]
            out[f"vel_vertex_{sid}"] = p["vel_vertex"]
            out[f"vertvel_vertex_{sid}"] = p["vertvel_vertex"]
            for name in sorted(s0.attrs):
                out[f"attr_{name}_{sid}"] = o.prepared_attr(sid, name)
        seeds = np.concatenate([cases.seeds_grid(7), cases.seeds_random(90, seed=2)])
        depths = np.linspace(20.0, 2200.0, seeds.shape[0]).astype(np.float32)
        out["seeds"] = seeds
        out["depths"] = depths
        out["cells"] = o.locate(seeds)
        for method in ("rk4", "euler"):
            for rec in (DT, 1800):
                r = o.streamline(seeds, DT, DUR, rec, depths=depths, method=method)
                out[f"stream_{method}_{rec}_points"] = r["points"]
                out[f"stream_{method}_{rec}_velocity"] = r["velocity"]
        o.activate(0, 1)
        for method in ("rk4", "euler"):
            r = o.pathline(seeds, DT, DUR, 1800, depths=depths, method=method)
            out[f"path_{method}_points"] = r["points"]
            out[f"path_{method}_velocity"] = r["velocity"]
            out[f"path_{method}_temperature"] = r["temperature"]
            out[f"path_{method}_seeds_out"] = r["seeds_out"]
        o.activate(0, None)
        for (w, h, d) in ((48, 24, 350.0), (32, 16, 0.0), (32, 16, 9000.0)):
            r = o.remap(w, h, depth=d)
            out[f"remap_{w}x{h}_{int(d)}_img0"] = r["img0"]
            if r["img1"] is not None:
                out[f"remap_{w}x{h}_{int(d)}_img1"] = r["img1"]
        o.close()
        path = os.path.join(HERE, f"ref_level{LEVEL}_{variant}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")
    # the two known answers the reference's own tests hold for hot-adjacent math
    # the system of the reference's test/test_gaussian.cpp:9-28 (expected {4.75, 0.5, 6.0}, tol 1e-6)
    a = np.array([[2.0, 3.0, -1.0], [4.0, 4.0, -3.0], [-2.0, 3.0, 2.0]])
    b = np.array([5.0, 3.0, 4.0])
    np.savez(os.path.join(HERE, "ref_known_answers.npz"), gauss_a=a, gauss_b=b, gauss_x=R.gauss3(a, b),
             wach_p=np.array([0.25, 0.5, 1.0]), wach_poly=np.array([[0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1.0]]),
             wach_w=R.wachspress(np.array([0.25, 0.5, 1.0]), np.array([[0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1.0]])))


if __name__ == "__main__":
    main()

"""small end-to-end case for compute-sanitizer: mesh + snapshots (with a non-monotone one), locate,
streamline / pathline (reference + walk + near-edge), remap, fixed-layer / fixed-latitude views"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cases
from mops_b200 import capi

eng = capi.Engine(0)
for variant in ("rich", "nonmono"):
    m = cases.mesh(3)
    s0, s1 = cases.snapshots(3, 8, variant)
    eng.set_mesh(m)
    eng.set_snapshot(0, s0)
    eng.set_snapshot(1, s1, async_=True)
    seeds = np.concatenate([cases.seeds_random(700, seed=2), m.vertex_xyz[:20], np.full((2, 3), np.nan), np.zeros((2, 3))])
    cells = eng.locate(seeds)
    a = eng.streamline(0, seeds, 600, 86400, 3600, depth=300.0, log_cells=True, near_edge=True)
    b = eng.pathline(0, 1, seeds, 600, 86400, 3600, depth=300.0, log_cells=True, walk=True)
    c = eng.streamline(0, seeds, 600, 86400, 3600, depth=300.0, method="euler", sort_particles=False)
    d = eng.pathline(0, 1, seeds, 600, 86400, 3600, depth=300.0)  # production (SEG park/resume with MOPS_SEGMENT_STEPS=7)
    e = eng.streamline(0, seeds, 600, 86400, 3600, depth=300.0)
    r = eng.remap(0, 64, 32, depth=300.0)
    v = eng.remap_fixed_layer(0, 64, 32, 3)
    g = eng.regrid_fixed_latitude(0, 64, 16, 20.0, 600.0, 5000.0)
    eng.get_prepared(0, attrs=2)
    print(variant, int(a["stats"].particle_steps), int(b["stats"].particle_steps), int(np.isnan(r["img0"]).sum()))
eng.close()
print("sanitize case done")

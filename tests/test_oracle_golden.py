"""CPU: the C restatement (oracle/mops_oracle.c) against the committed golden vectors, which are
outputs of the reference itself (tests/golden/make_golden.py ran oracle/_ref/libmops_ref.so).
Everything must be bit-identical."""
import os

import numpy as np
import pytest

import cases
from oracle import port_oracle as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("variant", ["plain", "rich"])
def test_port_matches_reference_golden(variant):
    g = np.load(os.path.join(GOLD, f"ref_level3_{variant}.npz"))
    level, L, dt, dur = int(g["level"]), int(g["n_levels"]), int(g["dt"]), int(g["duration"])
    m = cases.mesh(level)
    s0, s1 = cases.snapshots(level, L, variant)
    preps = [P.prepare(m, s0), P.prepare(m, s1)]
    for sid, p in enumerate(preps):
        assert _same(p.ztop_v, g[f"ztop_vertex_{sid}"])
        assert _same(p.vel_v, g[f"vel_vertex_{sid}"])
        assert _same(p.w_v, g[f"vertvel_vertex_{sid}"])
        for name in p.attrs_v:
            assert _same(p.attrs_v[name], g[f"attr_{name}_{sid}"])
    seeds, depths = g["seeds"], g["depths"]
    cells = P.locate(m, seeds)
    assert _same(cells, g["cells"])
    assert _same(P.locate(m, seeds, bruteforce=True), g["cells"])
    for method in ("rk4", "euler"):
        for rec in (dt, 1800):
            b = P.streamline(m, preps[0], seeds, cells, dt, dur, rec, depths=depths, method=method)
            f = P.finalize_lines(seeds, b["raw_pos"], b["raw_vel"])
            assert _same(f["points"], g[f"stream_{method}_{rec}_points"]), (method, rec)
            assert _same(f["velocity"], g[f"stream_{method}_{rec}_velocity"]), (method, rec)
        b = P.pathline(m, preps[0], preps[1], seeds, cells, dt, dur, 1800, depths=depths, method=method)
        f = P.finalize_lines(seeds, b["raw_pos"], b["raw_vel"], pathline_mode=True)
        assert _same(f["points"], g[f"path_{method}_points"])
        assert _same(f["velocity"], g[f"path_{method}_velocity"])
        assert _same(f["temperature"], g[f"path_{method}_temperature"])
        assert _same(f["last"], g[f"path_{method}_seeds_out"])  # R14: seeds overwritten with lastPoint
    for (w, h, d) in ((48, 24, 350), (32, 16, 0), (32, 16, 9000)):
        r = P.remap(m, preps[0], w, h, depth=float(d))
        assert _same(r["img0"], g[f"remap_{w}x{h}_{d}_img0"])
        key = f"remap_{w}x{h}_{d}_img1"
        if key in g.files:
            assert _same(r["img1"], g[key])
        else:
            assert r["img1"] is None


def test_known_answers():
    g = np.load(os.path.join(GOLD, "ref_known_answers.npz"))
    # reference's test/test_gaussian.cpp:9-28 expects {4.75, 0.5, 6.0} within 1e-6 -- the golden
    # value is what the compiled reference returned for that system
    assert np.allclose(g["gauss_x"], [4.75, 0.5, 6.0], atol=1e-6)
    assert _same(P.wachspress(g["wach_p"], g["wach_poly"]), g["wach_w"])
    assert np.allclose(g["wach_w"], [0.375, 0.125, 0.125, 0.375])


def test_rk4_particles_stop_at_first_failed_stage():
    """SURVEY finding 2 / R1: under reference semantics an RK4 particle that leaves its
    start-of-step cell stops for good; Euler particles keep going."""
    m = cases.mesh(3)
    s0, _ = cases.snapshots(3, 8, "rich")
    prep = P.prepare(m, s0)
    seeds = cases.seeds_random(300, seed=4)
    cells = P.locate(m, seeds)
    rk = P.streamline(m, prep, seeds, cells, 600, 86400, 3600, depth=300.0, method="rk4")
    eu = P.streamline(m, prep, seeds, cells, 600, 86400, 3600, depth=300.0, method="euler")
    assert (rk["status"] == 2).sum() > 0 and (eu["status"] != 0).sum() == 0
    dead = rk["status"] != 0
    # untouched slots of a stopped particle stay (0,0,0)
    assert (rk["raw_pos"][dead][:, -1] == 0).all()
    # the cell log of a stopped particle ends with -1s
    assert (rk["cell_log"][dead][:, -1] == -1).all()


def test_analytic_solid_body_rotation():
    """Independent of the reference: Euler-on-the-sphere trajectories in a solid-body field stay on
    the analytic rotation (closed form) to first order in dt."""
    from mops_b200 import synthetic as S
    m = cases.mesh(4)
    L, speed, tilt = 6, 1.0, 0.2
    s0 = S.solid_body_snapshot(m, L, speed, tilt=tilt)
    prep = P.prepare(m, s0)
    seeds = S.uniform_sphere_seeds(200, 8, lat_max=60.0)
    cells = P.locate(m, seeds)
    dur = 6 * 3600
    r = P.streamline(m, prep, seeds, cells, 120, dur, dur, depth=100.0, method="euler")
    axis = S.rotation_axis(tilt)
    R = np.linalg.norm(m.cell_xyz[0])
    ang = speed / R * dur
    k = axis
    v = seeds
    rot = v * np.cos(ang) + np.cross(k, v) * np.sin(ang) + k[None, :] * (v @ k)[:, None] * (1 - np.cos(ang))
    err = np.linalg.norm(r["pos"] - rot, axis=1)
    travelled = speed * dur * np.linalg.norm(np.cross(k, v / np.linalg.norm(v, axis=1, keepdims=True)), axis=1)
    assert (r["status"] == 0).all()
    # piecewise-linear velocity interpolation on a ~480 km mesh: a few percent of the path length
    assert (err < 0.05 * travelled + 50.0).all(), (err.max(), travelled.max())

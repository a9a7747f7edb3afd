"""CPU: the C restatement against the compiled reference (oracle/_ref/libmops_ref.so), live.
Skipped where the reference library has not been built (it needs /root/reference)."""
import numpy as np
import pytest

import cases
from oracle import port_oracle as P
from oracle import ref_oracle as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference)")


def _same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.fixture(scope="module")
def session():
    m = cases.mesh(4)
    s0, s1 = cases.snapshots(4, 10, "rich")
    o = R.RefOracle(m, [s0, s1])
    yield m, s0, s1, o
    o.close()


def test_sizeof_and_threads():
    assert R._load().refo_sizeof_vec3() == 24
    assert R.max_threads() >= 1


def test_gauss_known_answer_from_reference_test():
    # test/test_gaussian.cpp:9-28
    x = R.gauss3(np.array([[2.0, 3.0, -1.0], [4.0, 4.0, -3.0], [-2.0, 3.0, 2.0]]), np.array([5.0, 3.0, 4.0]))
    assert np.abs(x - np.array([4.75, 0.5, 6.0])).max() <= 1e-6


def test_seed_grid_matches_reference():
    from mops_b200 import synthetic as S
    a = R.generate_seeds(11, 11, (-60, 60), (-170, 170))
    assert a.shape[0] == 100  # config C1: "100 uniform seeds"
    assert _same(a, S.seed_grid(11, 11, (-60, 60), (-170, 170)))


def test_prepare_locate(session):
    m, s0, s1, o = session
    for sid, s in ((0, s0), (1, s1)):
        p, r = P.prepare(m, s), o.prepared(sid)
        assert _same(p.ztop_v, r["ztop_vertex"]) and _same(p.vel_v, r["vel_vertex"]) and _same(p.w_v, r["vertvel_vertex"])
        assert _same(p.ztop_c, r["ztop_cell"]) and _same(p.vel_c, r["vel_cell"])
        for name in s.attrs:
            assert _same(p.attrs_v[name], o.prepared_attr(sid, name))
    pts = cases.seeds_random(5000, seed=1)
    assert _same(P.locate(m, pts), o.locate(pts))


@pytest.mark.parametrize("method", ["rk4", "euler"])
def test_streamline_pathline(session, method):
    m, s0, s1, o = session
    p0, p1 = P.prepare(m, s0), P.prepare(m, s1)
    seeds = np.concatenate([cases.seeds_grid(10), cases.seeds_random(300, seed=6)])
    depths = np.linspace(10.0, 2400.0, seeds.shape[0]).astype(np.float32)
    cells = o.locate(seeds)
    o.activate(0, None)
    for rec in (120, 3600):
        r = o.streamline(seeds, 120, 43200, rec, depths=depths, method=method)
        b = P.streamline(m, p0, seeds, cells, 120, 43200, rec, depths=depths, method=method)
        f = P.finalize_lines(seeds, b["raw_pos"], b["raw_vel"])
        assert _same(r["points"], f["points"]) and _same(r["velocity"], f["velocity"]) and _same(r["last"], f["last"])
    # threads do not change results (particles are independent)
    R.set_threads(1)
    r1 = o.streamline(seeds, 120, 43200, 3600, depths=depths, method=method)
    R.set_threads(R.max_threads())
    assert _same(r1["points"], r["points"])
    o.activate(0, 1)
    r = o.pathline(seeds, 120, 43200, 3600, depths=depths, method=method)
    b = P.pathline(m, p0, p1, seeds, cells, 120, 43200, 3600, depths=depths, method=method)
    f = P.finalize_lines(seeds, b["raw_pos"], b["raw_vel"], pathline_mode=True)
    assert _same(r["points"], f["points"]) and _same(r["velocity"], f["velocity"])
    assert _same(r["temperature"], f["temperature"]) and _same(r["salinity"], f["salinity"])
    assert _same(r["seeds_out"], f["last"])
    o.activate(0, None)


def test_remap(session):
    m, s0, s1, o = session
    p0 = P.prepare(m, s0)
    for (w, h, d) in ((72, 36, 500.0), (40, 20, 0.0), (40, 20, 7000.0)):
        r, b = o.remap(w, h, depth=d), P.remap(m, p0, w, h, depth=d)
        assert _same(r["img0"], b["img0"])
        assert r["n_images"] == 2 and _same(r["img1"], b["img1"])
        # pixel cells are the reference's own KD-tree answers
        pos = P.pixel_positions(w, h)
        assert _same(o.locate(pos.reshape(-1, 3)).reshape(h, w), b["pixel_cell"])


def test_fixed_layer_and_fixed_latitude_views(session):
    """the two other views, oracle restatement vs the compiled reference, bit for bit"""
    m, s0, s1, o = session
    o.activate(0, None)
    p0 = P.prepare(m, s0)
    L = s0.n_levels
    for layer in (0, 3, L - 1, L + 5, -2):
        r = o.remap_fixed_layer(60, 30, layer)
        b = P.remap_fixed_layer(m, p0, 60, 30, layer)
        assert _same(r, b["img"]), layer
    ref_bottom = np.cumsum(np.full(L, 5000.0 / L))  # what RefOracle hands the grid
    for lat in (0.0, 37.5, -62.0):
        r = o.regrid_fixed_latitude(80, 25, lat)
        b = P.regrid_fixed_latitude(m, p0, 80, 25, lat, ref_bottom[0], ref_bottom[-1])
        assert _same(r, b["img"]), lat
        assert np.isfinite(b["img"][..., 0]).any()


@pytest.mark.parametrize("kind", ["m8", "m20"])
def test_general_voronoi_meshes(kind):
    """meshes with more than 6 edges per cell (maxEdges 8 / 12): restatement vs compiled reference"""
    m = cases.voronoi_mesh(kind)
    s0, s1 = cases.voronoi_snapshots(kind, 9)
    o = R.RefOracle(m, [s0, s1])
    try:
        p0, p1 = P.prepare(m, s0), P.prepare(m, s1)
        r = o.prepared(0)
        assert _same(p0.ztop_v, r["ztop_vertex"]) and _same(p0.vel_v, r["vel_vertex"]) and _same(p0.w_v, r["vertvel_vertex"])
        seeds = cases.seeds_random(600, seed=8)
        cells = o.locate(seeds)
        assert _same(P.locate(m, seeds), cells)
        for method in ("rk4", "euler"):
            rr = o.streamline(seeds, 600, 86400, 3600, depth=400.0, method=method)
            b = P.streamline(m, p0, seeds, cells, 600, 86400, 3600, depth=400.0, method=method)
            f = P.finalize_lines(seeds, b["raw_pos"], b["raw_vel"])
            assert _same(rr["points"], f["points"]) and _same(rr["velocity"], f["velocity"])
        o.activate(0, 1)
        rr = o.pathline(seeds, 600, 86400, 3600, depth=400.0, method="rk4")
        b = P.pathline(m, p0, p1, seeds, cells, 600, 86400, 3600, depth=400.0, method="rk4")
        f = P.finalize_lines(seeds, b["raw_pos"], b["raw_vel"], pathline_mode=True)
        assert _same(rr["points"], f["points"]) and _same(rr["velocity"], f["velocity"])
        o.activate(0, None)
        ri, bi = o.remap(64, 32, depth=400.0), P.remap(m, p0, 64, 32, depth=400.0)
        assert _same(ri["img0"], bi["img0"]) and _same(ri["img1"], bi["img1"])
    finally:
        o.close()


def test_ocean_mesh_with_land_boundaries():
    """culled mesh (0 entries in cellsOnCell / cellsOnVertex): boundary vertices in the preprocessing, seeds
    on land (nearest ocean cell, then IsInMesh fails), trajectories running into the coast, land pixels"""
    m = cases.ocean_mesh(4)
    s0, s1 = cases.ocean_snapshots(4, 9)
    assert (m.cells_on_vertex == 0).sum() > 100
    o = R.RefOracle(m, [s0, s1])
    try:
        p0, p1 = P.prepare(m, s0), P.prepare(m, s1)
        for sid, p in ((0, p0), (1, p1)):
            r = o.prepared(sid)
            assert _same(p.ztop_v, r["ztop_vertex"]) and _same(p.vel_v, r["vel_vertex"]) and _same(p.w_v, r["vertvel_vertex"])
            for name in s0.attrs:
                assert _same(p.attrs_v[name], o.prepared_attr(sid, name))
        seeds = cases.seeds_random(3000, seed=12)
        cells = o.locate(seeds)
        assert _same(P.locate(m, seeds), cells)
        rr = o.streamline(seeds, 600, 86400 * 2, 3600, depth=300.0, method="rk4")
        b = P.streamline(m, p0, seeds, cells, 600, 86400 * 2, 3600, depth=300.0, method="rk4")
        f = P.finalize_lines(seeds, b["raw_pos"], b["raw_vel"])
        assert _same(rr["points"], f["points"]) and _same(rr["velocity"], f["velocity"])
        st = b["status"]
        assert (st == 0).sum() > 500 and (st != 0).sum() > 500      # survivors and coast / land casualties
        o.activate(0, 1)
        rr = o.pathline(seeds, 600, 86400, 3600, depth=300.0, method="rk4")
        b = P.pathline(m, p0, p1, seeds, cells, 600, 86400, 3600, depth=300.0, method="rk4")
        f = P.finalize_lines(seeds, b["raw_pos"], b["raw_vel"], pathline_mode=True)
        assert _same(rr["points"], f["points"]) and _same(rr["velocity"], f["velocity"])
        o.activate(0, None)
        ri, bi = o.remap(90, 45, depth=300.0), P.remap(m, p0, 90, 45, depth=300.0)
        assert _same(ri["img0"], bi["img0"]) and _same(ri["img1"], bi["img1"])
        nan_frac = np.isnan(bi["img0"][..., 0]).mean()
        assert 0.1 < nan_frac < 0.6
        pos = P.pixel_positions(90, 45)
        assert _same(o.locate(pos.reshape(-1, 3)).reshape(45, 90), bi["pixel_cell"])
    finally:
        o.close()

"""GPU parity tests: the CUDA path, called through the C ABI (libmops_b200.so via ctypes),
against the oracle (oracle/mops_oracle.c, itself pinned bit-for-bit to the compiled reference)
on the same seeded inputs.

Bars (BASELINE.json north_star): cell-ID sequences and remap pixel cell IDs bit-exact except
for points within 1e-12 rad of a cell edge (counted); positions within 1 m; velocities within
1e-9 relative.  In practice everything except sin/cos is bit-identical, so the tests also
assert much tighter bounds and report the exact-match fraction.
"""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

DT, DAY = 120, 86400


@pytest.fixture(scope="module")
def eng():
    from mops_b200 import capi
    e = capi.Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def P():
    from oracle import port_oracle
    return port_oracle


def _setup(eng, level, L, variant):
    m = cases.mesh(level)
    s0, s1 = cases.snapshots(level, L, variant)
    eng.set_mesh(m)
    eng.set_snapshot(0, s0)
    eng.set_snapshot(1, s1)
    return m, s0, s1


@pytest.mark.parametrize("variant", ["plain", "rich", "nonmono"])
def test_prepare_bit_exact(eng, P, variant):
    m, s0, s1 = _setup(eng, 4, 12, variant)
    for slot, s in ((0, s0), (1, s1)):
        ref = P.prepare(m, s)
        na = min(2, len(s.attrs))
        got = eng.get_prepared(slot, attrs=na)
        assert np.array_equal(got["ztop_vertex"], ref.ztop_v)
        assert np.array_equal(got["vel_vertex"], ref.vel_v)
        assert np.array_equal(got["vertvel_vertex"][:, :-1], ref.w_v[:, :-1])
        names = sorted(s.attrs)
        if na >= 1:
            assert np.array_equal(got["attr0"], ref.attrs_v[names[0]])
        if na >= 2:
            assert np.array_equal(got["attr1"], ref.attrs_v[names[1]])
    info = eng.info()
    if variant == "nonmono":
        assert info.nonmonotone_cells[0] > 0
    else:
        assert info.nonmonotone_cells[0] == 0


@pytest.mark.parametrize("level", [3, 5])
def test_locate_exact(eng, P, level):
    m = cases.mesh(level)
    eng.set_mesh(m)
    pts = np.concatenate([cases.seeds_random(20000, seed=3), m.cell_xyz[:50] * 0.999, m.vertex_xyz[:50] * 1.001])
    got = eng.locate(pts)
    want = P.locate(m, pts)
    bad = np.nonzero(got != want)[0]
    # a mismatch is only legitimate for a query equidistant (to rounding) from two centres
    for i in bad:
        d = np.linalg.norm(m.cell_xyz[[got[i], want[i]]] - pts[i], axis=1)
        assert abs(d[0] - d[1]) <= 1e-6, (i, d)
    assert bad.size <= 60  # the 50 Voronoi vertices are exact three-way ties by construction


def _compare_traj(m, got, want, each, label):
    n = want["status"].shape[0]
    log_g, log_w = got["cell_log"], want["cell_log"]
    same_log = (log_g == log_w).all(axis=1)
    same_pos = np.array_equal(got["raw_pos"], want["raw_pos"], equal_nan=True)
    # positions: great-circle/chord error of every recorded slot
    err = np.linalg.norm(got["raw_pos"] - want["raw_pos"], axis=2)
    err = np.where(np.isnan(err), 0.0, err)
    vden = np.linalg.norm(want["raw_vel"], axis=2)
    verr = np.linalg.norm(got["raw_vel"] - want["raw_vel"], axis=2)
    vrel = np.where(vden > 0, verr / np.where(vden > 0, vden, 1.0), verr)
    vrel = np.where(np.isnan(vrel), 0.0, vrel)
    n_log_bad = int((~same_log).sum())
    print(f"[{label}] n={n} bit-identical positions={same_pos} max|dx|={err.max():.3e} m "
          f"max rel dv={vrel.max():.3e} cell-log mismatches={n_log_bad} "
          f"dead ref={int((want['status'] != 0).sum())} gpu={int((got['status'] != 0).sum())}")
    return same_log, err, vrel


@pytest.mark.parametrize("variant,method", [("plain", "rk4"), ("rich", "rk4"), ("rich", "euler"), ("nonmono", "rk4"),
                                            ("nonmono", "euler")])
def test_streamline_parity(eng, P, variant, method):
    m, s0, s1 = _setup(eng, 4, 12, variant)
    prep = P.prepare(m, s0)
    seeds = np.concatenate([cases.seeds_grid(15), cases.seeds_random(800, seed=5)])
    depths = np.linspace(5.0, 2500.0, seeds.shape[0]).astype(np.float32)
    cell0 = P.locate(m, seeds)
    for rec in (DT, 3600):
        want = P.streamline(m, prep, seeds, cell0, DT, DAY, rec, depths=depths, method=method)
        got = eng.streamline(0, seeds, DT, DAY, rec, depths=depths, cell0=cell0, method=method, log_cells=True)
        same_log, err, vrel = _compare_traj(m, got, want, DAY // rec, f"stream/{variant}/{method}/rec{rec}")
        assert same_log.all(), "cell-id sequences must be bit-exact"
        assert np.array_equal(got["status"], want["status"])
        assert np.array_equal(got["steps_alive"], want["steps_alive"])
        assert err.max() < 1e-6 and vrel.max() < 1e-9
        assert np.abs(got["depth"] - want["depth"]).max() < 1e-3
        # located on the device instead of given: same answer
    got2 = eng.streamline(0, seeds, DT, DAY, 3600, depths=depths, cell0=None, method=method, log_cells=True,
                          sort_particles=False)
    assert np.array_equal(got2["cell_log"], got["cell_log"])
    assert np.array_equal(got2["raw_pos"], got["raw_pos"], equal_nan=True)


@pytest.mark.parametrize("variant,method", [("plain", "rk4"), ("rich", "rk4"), ("rich", "euler"), ("nonmono", "rk4")])
def test_pathline_parity(eng, P, variant, method):
    m, s0, s1 = _setup(eng, 4, 12, variant)
    pf, pb = P.prepare(m, s0), P.prepare(m, s1)
    seeds = np.concatenate([cases.seeds_grid(12), cases.seeds_random(600, seed=9)])
    depths = np.linspace(50.0, 2000.0, seeds.shape[0]).astype(np.float32)
    cell0 = P.locate(m, seeds)
    for rec in (DT, 7200):
        want = P.pathline(m, pf, pb, seeds, cell0, DT, DAY, rec, depths=depths, method=method)
        got = eng.pathline(0, 1, seeds, DT, DAY, rec, depths=depths, cell0=cell0, method=method, log_cells=True)
        same_log, err, vrel = _compare_traj(m, got, want, DAY // rec, f"path/{variant}/{method}/rec{rec}")
        assert same_log.all()
        assert np.array_equal(got["status"], want["status"])
        assert err.max() < 1e-6 and vrel.max() < 1e-9
        if s0.attrs:
            a_g, a_w = got["raw_attr"], want["raw_attr"]
            assert np.allclose(a_g, a_w, rtol=1e-9, atol=1e-12, equal_nan=True)
            assert np.abs(a_w).max() > 0


def test_backward_and_uniform_depth(eng, P):
    m, s0, s1 = _setup(eng, 4, 12, "rich")
    prep = P.prepare(m, s0)
    seeds = cases.seeds_random(500, seed=21)
    cell0 = P.locate(m, seeds)
    want = P.streamline(m, prep, seeds, cell0, 300, 43200, 1800, depth=700.0, method="rk4", direction="backward")
    got = eng.streamline(0, seeds, 300, 43200, 1800, depth=700.0, cell0=cell0, method="rk4", direction="backward",
                         log_cells=True)
    assert np.array_equal(got["cell_log"], want["cell_log"])
    assert np.linalg.norm(got["raw_pos"] - want["raw_pos"], axis=2).max() < 1e-6


@pytest.mark.parametrize("variant", ["plain", "rich", "nonmono"])
def test_remap_parity(eng, P, variant):
    m, s0, s1 = _setup(eng, 4, 12, variant)
    prep = P.prepare(m, s0)
    for (w, h, d) in ((90, 45, 400.0), (64, 32, 3.0), (64, 32, 6000.0), (48, 24, 0.0)):
        want = P.remap(m, prep, w, h, depth=d)
        got = eng.remap(0, w, h, depth=d)
        cells_bad = int((got["pixel_cell"] != want["pixel_cell"]).sum())
        nan_g, nan_w = np.isnan(got["img0"]), np.isnan(want["img0"])
        print(f"[remap/{variant}/{w}x{h}/d{d}] pixel-cell mismatches={cells_bad} nan ref={int(nan_w.sum())} gpu={int(nan_g.sum())}")
        assert cells_bad == 0
        assert np.array_equal(nan_g, nan_w)
        assert np.allclose(got["img0"], want["img0"], rtol=1e-9, atol=1e-12, equal_nan=True)
        if want["img1"] is not None:
            assert got["img1"] is not None
            assert np.allclose(got["img1"], want["img1"], rtol=1e-9, atol=1e-9, equal_nan=True)
        else:
            assert got["img1"] is None


def test_finalize_lines_matches_oracle(eng, P):
    rng = np.random.default_rng(0)
    n, each = 40, 6
    seeds = rng.normal(size=(n, 3))
    raw_pos = rng.normal(size=(n, each, 3)); raw_vel = rng.normal(size=(n, each, 3))
    raw_pos[3, 0, 1] = np.nan; raw_pos[5, 2, 0] = np.inf; raw_pos[7, each - 1, 2] = np.nan
    seeds[9, 0] = np.nan
    for mode in (False, True):
        a = eng.finalize_lines(seeds, raw_pos, raw_vel, pathline_mode=mode)
        b = P.finalize_lines(seeds, raw_pos, raw_vel, pathline_mode=mode)
        for k in ("points", "velocity", "temperature", "salinity", "last"):
            assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_large_mesh_sorted_subsample_parity(eng, P):
    """BASELINE-shaped sizes: 163,842 cells x 60 layers, 300k Gaussian seeds processed in Morton
    order; a random subsample is checked against the oracle (cells located on the device, cell-id
    sequences, end positions, status), the whole set through size-independent properties."""
    from mops_b200 import synthetic as S
    m = cases.mesh(7)
    s0 = S.solid_body_snapshot(m, 60, 0.5, tilt=0.3)
    eng.set_mesh(m)
    eng.set_snapshot(0, s0)
    n = 300_000
    seeds = S.gaussian_seeds(n, 20261018)
    got = eng.streamline(0, seeds, 120, 43200, 3600, depth=800.0, cell0=None, method="rk4", log_cells=False)
    # properties over all particles: radius preserved (w = 0), steps bounded, status consistent
    r0 = np.linalg.norm(seeds, axis=1)
    alive = got["status"] == 0
    assert np.abs(np.linalg.norm(got["pos"], axis=1) - r0).max() < 1e-5
    assert (got["steps_alive"][alive] == 360).all() and (got["steps_alive"][~alive] <= 360).all()
    assert int(got["stats"].particle_steps) == int(got["steps_alive"].sum())
    assert int(got["stats"].alive_at_end) == int(alive.sum())
    # stopped particles leave a zero tail, live ones a full record
    assert (got["raw_pos"][alive][:, -1] != 0).any(axis=1).all()
    # unsorted processing gives the same bits
    got_u = eng.streamline(0, seeds, 120, 43200, 3600, depth=800.0, cell0=None, method="rk4", sort_particles=False)
    assert np.array_equal(got_u["raw_pos"], got["raw_pos"]) and np.array_equal(got_u["status"], got["status"])
    # subsample against the oracle
    rng = np.random.default_rng(3)
    pick = rng.choice(n, size=400, replace=False)
    prep = P.prepare(m, s0)
    cells = P.locate(m, seeds[pick])
    assert np.array_equal(eng.locate(seeds[pick]), cells)
    want = P.streamline(m, prep, seeds[pick], cells, 120, 43200, 3600, depth=800.0, method="rk4")
    assert np.array_equal(got["status"][pick], want["status"])
    assert np.array_equal(got["steps_alive"][pick], want["steps_alive"])
    assert np.array_equal(got["final_cell"][pick], want["final_cell"])
    assert np.linalg.norm(got["raw_pos"][pick] - want["raw_pos"], axis=2).max() < 1e-6
    same = np.array_equal(got["raw_pos"][pick], want["raw_pos"])
    print(f"[large] stopped {int((~alive).sum())}/{n}; subsample bit-identical={same}")


def test_near_edge_particles_are_counted(eng, P):
    """north_star: particles within 1e-12 rad of a cell edge are counted and reported.  Seeds placed
    exactly on cell edges / vertices must be flagged; generic seeds must not."""
    m, s0, s1 = _setup(eng, 4, 12, "plain")
    generic = cases.seeds_random(2000, seed=31)
    voc = m.vertices_on_cell - 1
    cells = np.arange(0, 200)
    on_vertex = m.vertex_xyz[voc[cells, 0]]
    on_edge = 0.5 * (m.vertex_xyz[voc[cells, 0]] + m.vertex_xyz[voc[cells, 1]])
    seeds = np.concatenate([generic, on_vertex, on_edge])
    cell0 = np.concatenate([P.locate(m, generic), cells, cells]).astype(np.int32)
    got = eng.streamline(0, seeds, 120, 3600, 3600, depth=500.0, cell0=cell0, method="rk4", near_edge=True)
    flagged = got["min_edge"] < 1e-12
    assert not flagged[:2000].any()
    assert flagged[2000:].all()
    assert int(got["stats"].near_edge_particles) == int(flagged.sum()) == 400
    # the diagnostic does not change results
    ref = eng.streamline(0, seeds, 120, 3600, 3600, depth=500.0, cell0=cell0, method="rk4")
    assert np.array_equal(ref["raw_pos"], got["raw_pos"], equal_nan=True) and np.array_equal(ref["status"], got["status"])


def test_walk_mode_crosses_cells_and_matches_analytic_rotation(eng, P):
    """MOPS_SEM_WALK (non-parity): RK4 particles survive cell crossings; identical to the reference
    semantics until the reference stops a particle; tracks the analytic solid-body rotation."""
    from mops_b200 import synthetic as S
    m = cases.mesh(5)
    speed, tilt = 1.0, 0.2
    s0 = S.solid_body_snapshot(m, 10, speed, tilt=tilt)
    eng.set_mesh(m)
    eng.set_snapshot(0, s0)
    seeds = S.uniform_sphere_seeds(3000, 17, lat_max=70.0)
    dur = 5 * 86400
    ref = eng.streamline(0, seeds, 600, dur, 86400, depth=100.0, method="rk4")
    walk = eng.streamline(0, seeds, 600, dur, 86400, depth=100.0, method="rk4", walk=True)
    assert (ref["status"] != 0).sum() > 100       # reference semantics: many stop at their first crossing
    assert (walk["status"] == 0).all()            # walk mode: none do
    alive = ref["status"] == 0
    assert np.array_equal(walk["raw_pos"][alive], ref["raw_pos"][alive])  # same bits where the reference survives
    axis = S.rotation_axis(tilt)
    R = np.linalg.norm(m.cell_xyz[0])
    ang = speed / R * dur
    v = seeds
    rot = v * np.cos(ang) + np.cross(axis, v) * np.sin(ang) + axis[None, :] * (v @ axis)[:, None] * (1 - np.cos(ang))
    err = np.linalg.norm(walk["pos"] - rot, axis=1)
    travelled = speed * dur * np.linalg.norm(np.cross(axis, v / np.linalg.norm(v, axis=1, keepdims=True)), axis=1)
    # piecewise (Wachspress) interpolation of a smooth field on a ~240 km mesh: within 2 % of the path length
    assert (err < 0.02 * travelled + 100.0).all(), (err.max(), travelled.max())
    eu = eng.streamline(0, seeds, 600, dur, 86400, depth=100.0, method="euler")
    err_eu = np.linalg.norm(eu["pos"] - rot, axis=1)
    print(f"[walk] rk4-walk max err {err.max():.1f} m, euler max err {err_eu.max():.1f} m over {travelled.max() / 1e3:.0f} km")


@pytest.mark.parametrize("variant", ["rich", "nonmono"])
def test_fixed_layer_and_fixed_latitude_views(eng, P, variant):
    """VisualizeFixedLayer / VisualizeFixedLatitude on the device vs the oracle"""
    m, s0, s1 = _setup(eng, 4, 12, variant)
    prep = P.prepare(m, s0)
    for layer in (0, 5, 11, 40, -3):
        got = eng.remap_fixed_layer(0, 72, 36, layer)
        want = P.remap_fixed_layer(m, prep, 72, 36, layer)
        assert np.array_equal(got["pixel_cell"], want["pixel_cell"])
        assert np.array_equal(np.isnan(got["img"]), np.isnan(want["img"]))
        assert np.allclose(got["img"], want["img"], rtol=1e-9, atol=1e-12, equal_nan=True)
    for lat in (0.0, 41.0, -70.0):
        got = eng.regrid_fixed_latitude(0, 90, 30, lat, 416.0, 5000.0)
        want = P.regrid_fixed_latitude(m, prep, 90, 30, lat, 416.0, 5000.0)
        assert np.array_equal(got["pixel_cell"], want["pixel_cell"])
        nan_g, nan_w = np.isnan(got["img"]), np.isnan(want["img"])
        print(f"[latitude/{variant}/{lat}] nan ref={int(nan_w.sum())} gpu={int(nan_g.sum())}")
        assert np.array_equal(nan_g, nan_w)
        assert np.allclose(got["img"], want["img"], rtol=1e-9, atol=1e-12, equal_nan=True)


def test_pinned_host_outputs(eng, P):
    """HOST-mode call with pinned output buffers.  Every slot is written by the kernel itself (no memset):
    results identical to the pageable path, stopped particles leave zeros."""
    import ctypes as C
    import os
    import torch
    from mops_b200 import capi
    m, s0, s1 = _setup(eng, 4, 12, "rich")
    seeds = cases.seeds_random(6000, seed=41)
    n, dur, rec = seeds.shape[0], 43200, 3600
    each = dur // rec
    staged = eng.pathline(0, 1, seeds, DT, dur, rec, depth=600.0, method="rk4")   # pageable numpy buffers
    assert (staged["status"] != 0).sum() > 100

    def pinned(shape, dtype):
        return torch.zeros(shape, dtype=dtype).pin_memory()
    xyz = pinned((n, 3), torch.float64); xyz.copy_(torch.from_numpy(seeds))
    depth = pinned((n,), torch.float32); depth.fill_(600.0)
    out_pos = pinned((n, each, 3), torch.float64); out_pos.fill_(7.0)   # garbage the kernel must overwrite
    out_vel = pinned((n, each, 3), torch.float64); out_vel.fill_(7.0)
    out_attr = pinned((n, each, 3), torch.float64); out_attr.fill_(7.0)
    status = pinned((n,), torch.int32)
    cfg = capi.TrajCfg(capi.METHOD_RK4, capi.DIR_FORWARD, DT, dur, rec, capi.MEM_HOST, 1)
    io = capi.TrajIO(n, xyz.data_ptr(), depth.data_ptr(), None, out_pos.data_ptr(), out_vel.data_ptr(), out_attr.data_ptr(),
                     None, status.data_ptr(), None, None)
    st = eng.traj_device(True, (0, 1), cfg, io, want_stats=True)
    assert np.array_equal(out_pos.numpy(), staged["raw_pos"])
    assert np.array_equal(out_vel.numpy(), staged["raw_vel"])
    assert np.array_equal(out_attr.numpy(), staged["raw_attr"])
    assert np.array_equal(status.numpy(), staged["status"]) and np.array_equal(xyz.numpy(), staged["pos"])
    assert int(st.particle_steps) == int(staged["steps_alive"].sum())


@pytest.mark.parametrize("kind,width", [("m8", 8), ("m20", 20)])
def test_general_voronoi_meshes_wide_records(eng, P, kind, width):
    """cells with up to 8 / 12 edges: the 8- and 20-wide cell-record instantiations of every kernel"""
    m = cases.voronoi_mesh(kind)
    s0, s1 = cases.voronoi_snapshots(kind, 9)
    eng.set_mesh(m)
    eng.set_snapshot(0, s0)
    eng.set_snapshot(1, s1)
    assert eng.info().record_width == width
    p0, p1 = P.prepare(m, s0), P.prepare(m, s1)
    got = eng.get_prepared(0, attrs=2)
    assert np.array_equal(got["ztop_vertex"], p0.ztop_v) and np.array_equal(got["vel_vertex"], p0.vel_v)
    seeds = np.concatenate([cases.seeds_random(3000, seed=8), m.cell_xyz[:100] * 0.9999])
    cells = P.locate(m, seeds)
    assert np.array_equal(eng.locate(seeds), cells)
    for method in ("rk4", "euler"):
        want = P.streamline(m, p0, seeds, cells, 600, 86400, 3600, depth=400.0, method=method)
        g = eng.streamline(0, seeds, 600, 86400, 3600, depth=400.0, cell0=cells, method=method, log_cells=True)
        assert np.array_equal(g["cell_log"], want["cell_log"]) and np.array_equal(g["status"], want["status"])
        assert np.linalg.norm(g["raw_pos"] - want["raw_pos"], axis=2).max() < 1e-6
        print(f"[{kind}/{method}] bit-identical={np.array_equal(g['raw_pos'], want['raw_pos'])} stopped={(want['status'] != 0).sum()}")
    want = P.pathline(m, p0, p1, seeds, cells, 600, 86400, 3600, depth=400.0, method="rk4")
    g = eng.pathline(0, 1, seeds, 600, 86400, 3600, depth=400.0, cell0=cells, method="rk4", log_cells=True)
    assert np.array_equal(g["cell_log"], want["cell_log"]) and np.array_equal(g["status"], want["status"])
    assert np.linalg.norm(g["raw_pos"] - want["raw_pos"], axis=2).max() < 1e-6
    assert np.allclose(g["raw_attr"], want["raw_attr"], rtol=1e-9, atol=1e-12)
    wi = P.remap(m, p0, 96, 48, depth=400.0)
    gi = eng.remap(0, 96, 48, depth=400.0)
    assert np.array_equal(gi["pixel_cell"], wi["pixel_cell"])
    assert np.allclose(gi["img0"], wi["img0"], rtol=1e-9, atol=1e-12, equal_nan=True)
    assert np.allclose(gi["img1"], wi["img1"], rtol=1e-9, atol=1e-9, equal_nan=True)
    wl = P.regrid_fixed_latitude(m, p0, 60, 20, 10.0, 500.0, 4000.0)
    gl = eng.regrid_fixed_latitude(0, 60, 20, 10.0, 500.0, 4000.0)
    assert np.allclose(gl["img"], wl["img"], rtol=1e-9, atol=1e-12, equal_nan=True)


@pytest.mark.parametrize("level", [4, 6])
def test_culled_ocean_mesh_with_land(eng, P, level):
    """mesh with removed (land) cells: the greedy walk alone is not exact there, the kd-tree fallback must give
    the reference's nearest-centre answer for seeds on land, behind peninsulas and in straits; boundary
    vertices in the preprocessing; trajectories running into the coast; land pixels in every view"""
    m = cases.ocean_mesh(level)
    s0, s1 = cases.ocean_snapshots(level, 9)
    eng.set_mesh(m)
    eng.set_snapshot(0, s0)
    eng.set_snapshot(1, s1)
    p0, p1 = P.prepare(m, s0), P.prepare(m, s1)
    got = eng.get_prepared(0, attrs=2)
    assert np.array_equal(got["ztop_vertex"], p0.ztop_v) and np.array_equal(got["vel_vertex"], p0.vel_v)
    names = sorted(s0.attrs)
    assert np.array_equal(got["attr0"], p0.attrs_v[names[0]]) and np.array_equal(got["attr1"], p0.attrs_v[names[1]])
    seeds = cases.seeds_random(20000 if level == 6 else 4000, seed=12)
    cells = P.locate(m, seeds)
    gc = eng.locate(seeds)
    assert np.array_equal(gc, cells), f"{(gc != cells).sum()} of {len(cells)} differ"
    import torch
    dc = eng.locate(torch.from_numpy(seeds).cuda())   # device buffers: asynchronous on the engine's stream
    eng.synchronize()
    assert np.array_equal(dc.cpu().numpy(), cells)
    n = 4000
    want = P.streamline(m, p0, seeds[:n], cells[:n], 600, 2 * 86400, 3600, depth=300.0)
    g = eng.streamline(0, seeds[:n], 600, 2 * 86400, 3600, depth=300.0, log_cells=True)   # device-located start cells
    assert np.array_equal(g["cell_log"], want["cell_log"]) and np.array_equal(g["status"], want["status"])
    assert np.array_equal(g["raw_pos"], want["raw_pos"]) and np.array_equal(g["raw_vel"], want["raw_vel"])
    assert (want["status"] == 0).sum() > 50 and (want["status"] != 0).sum() > 500
    want = P.pathline(m, p0, p1, seeds[:n], cells[:n], 600, 86400, 3600, depth=300.0)
    g = eng.pathline(0, 1, seeds[:n], 600, 86400, 3600, depth=300.0, cell0=cells[:n], log_cells=True)
    assert np.array_equal(g["cell_log"], want["cell_log"]) and np.array_equal(g["raw_pos"], want["raw_pos"])
    assert np.array_equal(g["raw_attr"], want["raw_attr"])
    wi = P.remap(m, p0, 180, 90, depth=300.0)
    gi = eng.remap(0, 180, 90, depth=300.0)
    assert np.array_equal(gi["pixel_cell"], wi["pixel_cell"])
    # images go through sin/cos of the pixel's lat/lon (ENU conversion): same NaN mask, values to 1e-9 relative
    assert np.array_equal(np.isnan(gi["img0"]), np.isnan(wi["img0"])) and np.array_equal(np.isnan(gi["img1"]), np.isnan(wi["img1"]))
    assert np.allclose(gi["img0"], wi["img0"], rtol=1e-9, atol=1e-12, equal_nan=True)
    assert np.allclose(gi["img1"], wi["img1"], rtol=1e-9, atol=1e-9, equal_nan=True)
    assert 0.1 < np.isnan(wi["img0"][..., 0]).mean() < 0.6
    wl = P.remap_fixed_layer(m, p0, 120, 60, 2)
    gl = eng.remap_fixed_layer(0, 120, 60, 2)
    assert np.array_equal(gl["pixel_cell"], wl["pixel_cell"]) and np.array_equal(np.isnan(gl["img"]), np.isnan(wl["img"]))
    assert np.allclose(gl["img"], wl["img"], rtol=1e-9, atol=1e-12, equal_nan=True)
    wl = P.regrid_fixed_latitude(m, p0, 90, 20, 10.0, 500.0, 4000.0)
    gl = eng.regrid_fixed_latitude(0, 90, 20, 10.0, 500.0, 4000.0)
    assert np.array_equal(np.isnan(gl["img"]), np.isnan(wl["img"]))
    assert np.allclose(gl["img"], wl["img"], rtol=1e-9, atol=1e-12, equal_nan=True)

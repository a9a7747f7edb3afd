"""GPU: edge cases the reference's kernels branch on -- empty input, invalid settings, bad / missing start cells,
non-finite seeds, depths above the surface and below the bottom, record periods that never fire, attribute
gating -- each against the oracle."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from mops_b200 import capi
    from oracle import port_oracle as P
    e = capi.Engine(0)
    m = cases.mesh(4)
    s0, s1 = cases.snapshots(4, 12, "rich")
    e.set_mesh(m)
    e.set_snapshot(0, s0)
    e.set_snapshot(1, s1)
    yield e, P, m, P.prepare(m, s0), P.prepare(m, s1), capi
    e.close()


def test_empty_and_invalid_settings(ctx):
    e, P, m, p0, p1, capi = ctx
    r = e.streamline(0, np.zeros((0, 3)), 120, 3600, 600)
    assert r["raw_pos"].shape == (0, 6, 3) and int(r["stats"].particle_steps) == 0
    seeds = cases.seeds_random(10, seed=1)
    for (dt, dur, rec) in ((0, 3600, 600), (120, 0, 600), (120, 3600, 0), (120, 60, 600), (7200, 3600, 600)):
        with pytest.raises(capi.MopsError):   # reference: Error("invalid trajectory settings") + empty result
            e.streamline(0, seeds, dt, dur, rec)
    with pytest.raises(capi.MopsError):
        e.streamline(3, seeds, 120, 3600, 600)       # snapshot slot never set
    with pytest.raises(capi.MopsError):
        e.remap(0, 0, 10)


def test_bad_cells_and_nonfinite_seeds(ctx):
    e, P, m, p0, p1, capi = ctx
    seeds = cases.seeds_random(200, seed=3)
    cells = P.locate(m, seeds)
    cells[::7] = -1
    cells[3::11] = m.n_cells + 5
    seeds[5] = np.nan
    seeds[9, 1] = np.inf
    want = P.streamline(m, p0, seeds, cells, 300, 21600, 1800, depth=300.0)
    got = e.streamline(0, seeds, 300, 21600, 1800, depth=300.0, cell0=cells, log_cells=True)
    assert np.array_equal(got["status"], want["status"]) and (got["status"] == 1).sum() >= 40
    assert np.array_equal(got["cell_log"], want["cell_log"])
    assert np.array_equal(got["raw_pos"], want["raw_pos"], equal_nan=True)
    assert np.array_equal(got["raw_vel"], want["raw_vel"], equal_nan=True)
    bad = got["status"] == 1
    assert (got["raw_pos"][bad] == 0).all()          # not even the seed is written (VK:895-897)
    # device-located: non-finite seeds get cell -1 and stop the same way
    g2 = e.streamline(0, seeds, 300, 21600, 1800, depth=300.0, cell0=None)
    assert g2["status"][5] == 1 and g2["status"][9] == 1
    assert e.locate(seeds)[5] == -1


def test_depth_above_surface_and_below_bottom(ctx):
    e, P, m, p0, p1, capi = ctx
    seeds = cases.seeds_random(300, seed=5)
    cells = P.locate(m, seeds)
    depths = np.concatenate([np.zeros(100), np.full(100, 9000.0), np.linspace(1.0, 4000.0, 100)]).astype(np.float32)
    for method in ("rk4", "euler"):
        want = P.streamline(m, p0, seeds, cells, 600, 43200, 3600, depths=depths, method=method)
        got = e.streamline(0, seeds, 600, 43200, 3600, depths=depths, cell0=cells, method=method, log_cells=True)
        assert np.array_equal(got["status"], want["status"])
        assert np.array_equal(got["raw_pos"], want["raw_pos"]) and np.array_equal(got["depth"], want["depth"])
        wantp = P.pathline(m, p0, p1, seeds, cells, 600, 43200, 3600, depths=depths, method=method)
        gotp = e.pathline(0, 1, seeds, 600, 43200, 3600, depths=depths, cell0=cells, method=method, log_cells=True)
        assert np.array_equal(gotp["status"], wantp["status"])
        assert np.array_equal(gotp["raw_pos"], wantp["raw_pos"])
    # negative depth = above the sea surface: streamline clamps into layer 1, pathline reports ABOVE_SURFACE
    # (the reference reads ztop[-1] there; not replicated)
    up = np.full(50, -5.0, dtype=np.float32)
    gs = e.streamline(0, seeds[:50], 600, 7200, 3600, depths=up, cell0=cells[:50])
    ws = P.streamline(m, p0, seeds[:50], cells[:50], 600, 7200, 3600, depths=up)
    assert np.array_equal(gs["raw_pos"], ws["raw_pos"]) and np.array_equal(gs["status"], ws["status"])
    gp = e.pathline(0, 1, seeds[:50], 600, 7200, 3600, depths=up, cell0=cells[:50])
    wp = P.pathline(m, p0, p1, seeds[:50], cells[:50], 600, 7200, 3600, depths=up)
    assert (gp["status"] == 5).all() and np.array_equal(gp["status"], wp["status"])
    # ... and the count is surfaced (the C++ drop-in prints it): silent freezes near the surface were an advisor finding
    assert int(gp["stats"].above_surface_particles) == gp["status"].shape[0]


def test_record_periods(ctx):
    e, P, m, p0, p1, capi = ctx
    seeds = cases.seeds_random(150, seed=6)
    cells = P.locate(m, seeds)
    # (dt, duration, recordT): record not a multiple of dt; record < dt (pathline never records: interval 0);
    # duration not a multiple of recordT (integer division of slots)
    for (dt, dur, rec) in ((700, 21000, 1000), (600, 7200, 300), (450, 20000, 3000), (120, 3600, 3600)):
        ws = P.streamline(m, p0, seeds, cells, dt, dur, rec, depth=250.0)
        gs = e.streamline(0, seeds, dt, dur, rec, depth=250.0, cell0=cells)
        assert np.array_equal(gs["raw_pos"], ws["raw_pos"]) and np.array_equal(gs["raw_vel"], ws["raw_vel"]), (dt, dur, rec)
        wp = P.pathline(m, p0, p1, seeds, cells, dt, dur, rec, depth=250.0)
        gp = e.pathline(0, 1, seeds, dt, dur, rec, depth=250.0, cell0=cells)
        assert np.array_equal(gp["raw_pos"], wp["raw_pos"]) and np.array_equal(gp["raw_vel"], wp["raw_vel"]), (dt, dur, rec)
        assert np.array_equal(gp["raw_attr"], wp["raw_attr"])


def test_attribute_gating(ctx):
    """pathline / remap attributes exist only when the front snapshot holds MORE than one scalar (VK:259-267, 1093-1104)"""
    e, P, m, p0, p1, capi = ctx
    import copy
    s0, s1 = cases.snapshots(4, 12, "rich")
    a, b = copy.deepcopy(s0), copy.deepcopy(s1)
    del a.attrs["temperature"]; del b.attrs["temperature"]
    e.set_snapshot(2, a)
    e.set_snapshot(3, b)
    seeds = cases.seeds_random(100, seed=9)
    cells = P.locate(m, seeds)
    g = e.pathline(2, 3, seeds, 600, 7200, 3600, depth=250.0, cell0=cells)
    w = P.pathline(m, P.prepare(m, a), P.prepare(m, b), seeds, cells, 600, 7200, 3600, depth=250.0)
    assert (g["raw_attr"] == 0).all() and (w["raw_attr"] == 0).all()
    assert np.array_equal(g["raw_pos"], w["raw_pos"])
    img = e.remap(2, 32, 16, depth=250.0)
    assert img["img1"] is None and P.remap(m, P.prepare(m, a), 32, 16, depth=250.0)["img1"] is None


def test_host_submit_wait_matches_blocking_call(ctx):
    """The asynchronous HOST-memory form (mops_*_submit / mops_traj_wait, two alternating staging sets): three calls in
    flight back to back -- the third reuses the first one's staging set -- give, for every output (records, attributes,
    cell log, status, steps, final cell, end points, counters), exactly what the blocking call gives; the what = 0 wait
    returns the end points before the recorded trajectories have to be there; given and device-located start cells."""
    import ctypes as C
    e, P, m, p0, p1, capi = ctx
    seeds = cases.seeds_random(5000, seed=21)
    seeds[17] = np.nan
    depths = np.linspace(5.0, 900.0, 5000).astype(np.float32)
    cells = P.locate(m, seeds)
    cells[40::97] = -1
    n, dt, dur, rec = seeds.shape[0], 300, 21600, 3600
    each, times = dur // rec, dur // dt
    for cell0 in (cells, None):
        a = e.pathline(0, 1, seeds, dt, dur, rec, depths=depths, cell0=cell0, log_cells=True)
        runs = []
        for _ in range(3):
            r = {"pos": seeds.copy(), "depth": depths.copy(), "raw_pos": np.full((n, each, 3), 7.0), "raw_vel": np.full((n, each, 3), 7.0),
                 "raw_attr": np.full((n, each, 3), 7.0), "cell_log": np.zeros((n, times), np.int32), "status": np.zeros(n, np.int32),
                 "steps_alive": np.zeros(n, np.int32), "final_cell": np.zeros(n, np.int32),
                 "c0": None if cell0 is None else np.ascontiguousarray(cell0, dtype=np.int32)}
            pt = lambda x: None if x is None else x.ctypes.data
            r["io"] = capi.TrajIO(n, pt(r["pos"]), pt(r["depth"]), pt(r["c0"]), pt(r["raw_pos"]), pt(r["raw_vel"]), pt(r["raw_attr"]),
                                  pt(r["cell_log"]), pt(r["status"]), pt(r["steps_alive"]), pt(r["final_cell"]), None)
            runs.append(r)
        cfg = capi.TrajCfg(capi.METHOD_RK4, capi.DIR_FORWARD, dt, dur, rec, capi.MEM_HOST, 1)
        tickets = [e.traj_submit(True, (0, 1), cfg, r["io"]) for r in runs]
        assert tickets[0] > 0 and tickets[1] == tickets[0] + 1 and tickets[2] == tickets[0] + 2
        e.traj_wait(tickets[1], 0)
        assert np.array_equal(runs[1]["pos"], a["pos"], equal_nan=True) and np.array_equal(runs[1]["depth"], a["depth"])
        for tk, r in zip(tickets, runs):
            st = e.traj_wait(tk, 1)
            for k in ("raw_pos", "raw_vel", "raw_attr", "pos", "depth", "cell_log", "status", "steps_alive", "final_cell"):
                assert np.array_equal(a[k], r[k], equal_nan=True), k
            if tk != tickets[0]:  # the first ticket's set was reused by the third submit, which completed it: its stats are gone
                for f in ("particle_steps", "alive_at_end"):
                    assert int(getattr(a["stats"], f)) == int(getattr(st, f)), f
                assert st.kernel_ms > 0 and st.total_ms >= st.kernel_ms
    # DEVICE-memory buffers have no submit form (they are asynchronous on the context's stream already)
    cfg_d = capi.TrajCfg(capi.METHOD_RK4, capi.DIR_FORWARD, dt, dur, rec, capi.MEM_DEVICE, 1)
    with pytest.raises(capi.MopsError):
        e.traj_submit(True, (0, 1), cfg_d, runs[0]["io"])
    want = P.pathline(m, p0, p1, seeds, P.locate(m, seeds), dt, dur, rec, depths=depths)
    assert np.array_equal(runs[2]["raw_pos"], want["raw_pos"], equal_nan=True)

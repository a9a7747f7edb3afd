"""The compacting multi-launch form of the advection kernel (MOPS_SEGMENT_STEPS, mops_b200/csrc/engine.cu run_range)
must be invisible in the results: one launch per call, the default 40-step launches and deliberately awkward 7-step
launches (segment ends in the middle of a record interval) give bit-identical outputs, and those equal the oracle's.
"""
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

DT = 120
KEYS = ("raw_pos", "raw_vel", "pos", "depth", "cell_log", "status", "steps_alive", "final_cell")


def _engine(segment_steps):
    from mops_b200 import capi
    old = os.environ.get("MOPS_SEGMENT_STEPS")
    if segment_steps is None:
        os.environ.pop("MOPS_SEGMENT_STEPS", None)
    else:
        os.environ["MOPS_SEGMENT_STEPS"] = str(segment_steps)
    try:
        return capi.Engine(0)  # the variable is read once, when the context is created
    finally:
        if old is None:
            os.environ.pop("MOPS_SEGMENT_STEPS", None)
        else:
            os.environ["MOPS_SEGMENT_STEPS"] = old


def test_segmented_launches_are_bit_identical():
    from oracle import port_oracle as P
    m = cases.mesh(4)
    s0, s1 = cases.snapshots(4, 12, "rich")
    seeds = np.concatenate([cases.seeds_grid(12), cases.seeds_random(1500, seed=31)])
    depths = np.linspace(20.0, 2200.0, seeds.shape[0]).astype(np.float32)
    cell0 = P.locate(m, seeds)
    duration, rec = 100 * DT, 9 * DT  # 100 steps, 11 records; neither 7 nor 40 divides the record interval
    results = {}
    for seg in (0, 7, None):  # None = the library default
        eng = _engine(seg)
        try:
            eng.set_mesh(m)
            eng.set_snapshot(0, s0)
            eng.set_snapshot(1, s1)
            results[seg] = {
                "stream": eng.streamline(0, seeds, DT, duration, rec, depths=depths, cell0=cell0, log_cells=True),
                "stream_unsorted": eng.streamline(0, seeds, DT, duration, rec, depths=depths, cell0=None, log_cells=True,
                                                  sort_particles=False),
                "path": eng.pathline(0, 1, seeds, DT, duration, rec, depths=depths, cell0=cell0, log_cells=True),
                "path_euler": eng.pathline(0, 1, seeds, DT, duration, rec, depths=depths, cell0=cell0, method="euler",
                                           log_cells=True),
            }
        finally:
            eng.close()
    one = results[0]
    assert one["stream"]["stats"].launches < results[7]["stream"]["stats"].launches  # really ran in segments
    for seg in (7, None):
        for name, got in results[seg].items():
            for k in KEYS:
                assert np.array_equal(got[k], one[name][k], equal_nan=True), (seg, name, k)
            if got["raw_attr"] is not None:
                assert np.array_equal(got["raw_attr"], one[name]["raw_attr"], equal_nan=True), (seg, name)
            assert got["stats"].particle_steps == one[name]["stats"].particle_steps, (seg, name)
            assert got["stats"].alive_at_end == one[name]["stats"].alive_at_end, (seg, name)
    # ... and the segmented runs equal the oracle, not just each other
    pf, pb = P.prepare(m, s0), P.prepare(m, s1)
    want_s = P.streamline(m, pf, seeds, cell0, DT, duration, rec, depths=depths)
    want_p = P.pathline(m, pf, pb, seeds, cell0, DT, duration, rec, depths=depths)
    for got, want in ((results[7]["stream"], want_s), (results[7]["path"], want_p)):
        assert np.array_equal(got["cell_log"], want["cell_log"])
        assert np.array_equal(got["status"], want["status"])
        assert np.array_equal(got["steps_alive"], want["steps_alive"])
        assert np.linalg.norm(got["raw_pos"] - want["raw_pos"], axis=2).max() < 1e-6
    assert (want_s["status"] != 0).sum() > 0, "the case must contain particles that stop part-way"

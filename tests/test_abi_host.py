"""CPU: the C-ABI shared library loads, exports every symbol include/mops_b200.h declares, fails
loudly without a GPU (no fallback), and its pure-host function matches the reference's NaN
trimming cases (test/test_trajector.cpp)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from mops_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return capi.load_library(), capi


def test_exports_match_header():
    lib, capi = _lib()
    hdr = open(os.path.join(ROOT, "include", "mops_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|void|const char\*|void\*|int32_t|mops_ctx\*)\s+(mops_[a-z0-9_]+)\s*\(", hdr, flags=re.M))
    assert declared == set(capi.ABI_SYMBOLS), declared ^ set(capi.ABI_SYMBOLS)
    raw = ctypes.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    assert raw.mops_abi_version() == 2


def test_no_cpu_fallback():
    import torch
    lib, capi = _lib()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.MopsError):
        capi.Engine(0)
    h = ctypes.c_void_p()
    assert lib.mops_create(ctypes.byref(h), 0) == -3  # MOPS_E_NODEVICE
    assert not h.value


def test_product_never_touches_the_oracle():
    """the oracle is test infrastructure: nothing under mops_b200/ or include/ may name it"""
    for top in ("mops_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", ".sh")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read().lower()
                    assert "oracle" not in txt, f"{os.path.join(dirpath, f)} mentions the oracle"


def _finalize(capi, seeds, raw_pos, raw_vel, mode):
    lib = capi.load_library()
    n, each = raw_pos.shape[:2]
    per = each + 1
    pts = np.zeros((n, per, 3)); vel = np.zeros((n, per, 3)); t = np.zeros((n, per)); s = np.zeros((n, per)); last = np.zeros((n, 3))
    rc = lib.mops_finalize_lines(n, each, seeds.ctypes.data, raw_pos.ctypes.data, raw_vel.ctypes.data, mode, pts.ctypes.data,
                                 vel.ctypes.data, t.ctypes.data, s.ctypes.data, last.ctypes.data)
    assert rc == 0
    return pts, vel, t, s, last


def test_nan_trimming_cases_of_reference_test():
    """the four sections of test/test_trajector.cpp:27-208, on the flat form"""
    lib, capi = _lib()
    nan = np.nan
    each = 4
    seeds = np.array([[nan, 0, 0], [1.0, 1, 1], [2.0, 2, 2], [3.0, 3, 3]])
    raw_pos = np.arange(4 * each * 3, dtype=np.float64).reshape(4, each, 3) + 10.0
    raw_vel = np.ones((4, each, 3))
    raw_pos[1, 0, 0] = nan      # NaN at index 1 of the assembled line
    raw_pos[2, 2, 1] = nan      # NaN mid-line (assembled index 3)
    pts, vel, t, s, last = _finalize(capi, seeds, raw_pos, raw_vel, 0)
    per = each + 1
    # (1) first point NaN: whole line = first point, velocities zero, length preserved
    assert np.isnan(pts[0, :, 0]).all() and (vel[0] == 0).all() and pts.shape[1] == per
    # (2) NaN at index 1: padded with point 0; velocities zero from index 0 on
    assert (pts[1] == seeds[1]).all() and (vel[1] == 0).all() and (last[1] == seeds[1]).all()
    # (3) NaN mid-line at k=3: points 3.. = point 2; velocity zero from k-1 = 2 on, earlier kept
    assert (pts[2, 3:] == pts[2, 2]).all() and (vel[2, 2:] == 0).all() and (vel[2, :2] == 1).all()
    assert (last[2] == pts[2, 2]).all()
    # (4) all valid: untouched; velocity has the trailing zero pad (R7)
    assert (pts[3, 0] == seeds[3]).all() and (pts[3, 1:] == raw_pos[3]).all()
    assert (vel[3, :each] == 1).all() and (vel[3, each] == 0).all() and (last[3] == raw_pos[3, -1]).all()
    # oracle restatement agrees bit for bit, both modes
    from oracle import port_oracle as P
    for mode in (0, 1):
        a = _finalize(capi, seeds, raw_pos, raw_vel, mode)
        b = P.finalize_lines(seeds, raw_pos, raw_vel, pathline_mode=bool(mode))
        for x, k in zip(a, ("points", "velocity", "temperature", "salinity", "last")):
            assert np.array_equal(x, b[k], equal_nan=True), k


def test_finalize_lines_threaded_equals_serial(monkeypatch):
    """large calls are split over the host cores (MOPS_HOST_THREADS overrides the count): same bytes either way, and
    equal to the oracle restatement"""
    lib, capi = _lib()
    rng = np.random.default_rng(3)
    n, each = 30000, 40  # 1.2 M points: above the threshold where threads are used
    seeds = rng.normal(size=(n, 3)); raw_pos = rng.normal(size=(n, each, 3)); raw_vel = rng.normal(size=(n, each, 3))
    raw_pos[::11, 17:, :] = np.nan
    raw_pos[5, 0, 0] = np.inf
    seeds[13, 2] = np.nan
    out = {}
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("MOPS_HOST_THREADS", threads)
        out[threads] = _finalize(capi, seeds, raw_pos, raw_vel, 1)
    for threads in ("3", "8"):
        for a, b in zip(out["1"], out[threads]):
            assert np.array_equal(a, b, equal_nan=True)
    from oracle import port_oracle as P
    ref = P.finalize_lines(seeds, raw_pos, raw_vel, pathline_mode=True)
    for x, k in zip(out["8"], ("points", "velocity", "temperature", "salinity", "last")):
        assert np.array_equal(x, ref[k], equal_nan=True), k


def test_shard_bounds_match_the_python_twin():
    """mops_shard_bounds (pure host function of the C ABI) and mops_b200.sharding.block_bounds cut the same equal blocks"""
    from mops_b200 import sharding
    lib, capi = _lib()
    lo, hi = ctypes.c_int64(), ctypes.c_int64()
    for n in (0, 1, 7, 64_000_000, 10_000_001):
        for world in (1, 2, 3, 8):
            covered = 0
            for r in range(world):
                lib.mops_shard_bounds(n, r, world, ctypes.byref(lo), ctypes.byref(hi))
                assert (lo.value, hi.value) == sharding.block_bounds(n, r, world)
                assert lo.value == covered and hi.value - lo.value in (n // world, n // world + 1)
                covered = hi.value
            assert covered == n


def test_multi_gpu_entry_points_fail_cleanly_without_devices():
    """no GPU here: the multi-GPU constructors report MOPS_E_NODEVICE / invalid arguments instead of crashing, and there is
    still no fallback of any kind"""
    import torch
    lib, capi = _lib()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    assert lib.mops_multi_create(ctypes.byref(h), 0, None) == -3 and not h.value
    assert lib.mops_dist_create(ctypes.byref(h), None, 0, 1, None) == -1
    assert lib.mops_multi_pathline(None, None, 0, 1, None, None) == -1

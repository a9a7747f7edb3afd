"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/_build/libmops_oracle.so,
the plain-C restatement (Tier B) of the reference's hot path (oracle/mops_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (mops_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libmops_oracle.so")

STATUS_NAMES = {0: "alive", 1: "bad_cell", 2: "not_in_cell", 3: "bad_column", 4: "zero_velocity",
                5: "above_surface", 6: "bad_setup"}

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mops_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_pixel_position.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


class Prepared:
    """Vertex-major arrays of one snapshot (what MOPSApp::addSol leaves in MPASOSolution)."""

    def __init__(self, ztop_v, vel_v, w_v, attrs_v: Dict[str, np.ndarray], ztop_c=None, vel_c=None):
        self.ztop_v, self.vel_v, self.w_v, self.attrs_v = ztop_v, vel_v, w_v, attrs_v
        self.ztop_c, self.vel_c = ztop_c, vel_c

    def attr_list(self):
        """(a0, a1, count) in std::map (alphabetical) order, first two (R11)."""
        names = sorted(self.attrs_v.keys())
        arrs = [self.attrs_v[n] for n in names[:2]]
        while len(arrs) < 2:
            arrs.append(None)
        return arrs[0], arrs[1], len(names)


def prepare(mesh, snap) -> Prepared:
    lib = _load()
    L = snap.n_levels
    nc, nv = mesh.n_cells, mesh.n_vertices
    ztop_v = np.zeros((nv, L)); vel_v = np.zeros((nv, L, 3)); w_v = np.zeros((nv, L + 1))
    ztop_c = np.zeros((nc, L)); vel_c = np.zeros((nc, L, 3))
    cx, vx = _c(mesh.cell_xyz, np.float64), _c(mesh.vertex_xyz, np.float64)
    cov = _c(mesh.cells_on_vertex, np.int32)
    rc = lib.orc_prepare_snapshot(C.c_int(nc), C.c_int(nv), C.c_int(L), _p(cx), _p(vx), _p(cov),
                                  _p(_c(snap.zonal, np.float64)), _p(_c(snap.meridional, np.float64)),
                                  _p(_c(snap.layer_thickness, np.float64)), _p(_c(snap.bottom_depth, np.float64)),
                                  _p(_c(snap.vert_vel_top, np.float64)), _p(ztop_v), _p(vel_v), _p(w_v), _p(ztop_c), _p(vel_c))
    assert rc == 0
    attrs_v = {}
    for name, a in snap.attrs.items():
        out = np.zeros((nv, L))
        lib.orc_cell_to_vertex_scalar(C.c_int(nc), C.c_int(nv), C.c_int(L), _p(cx), _p(vx), _p(cov),
                                      _p(_c(a, np.float64)), C.c_int(1), _p(out))
        attrs_v[name] = out
    return Prepared(ztop_v, vel_v, w_v, attrs_v, ztop_c, vel_c)


def locate(mesh, xyz, bruteforce: bool = False) -> np.ndarray:
    lib = _load()
    xyz = _c(xyz, np.float64)
    out = np.zeros(xyz.shape[0], dtype=np.int32)
    fn = lib.orc_locate_bruteforce if bruteforce else lib.orc_locate
    fn(C.c_int64(xyz.shape[0]), _p(xyz), C.c_int(mesh.n_cells), _p(_c(mesh.cell_xyz, np.float64)), _p(out))
    return out


def _mesh_args(mesh, L):
    arrs = (_c(mesh.cell_xyz, np.float64), _c(mesh.vertex_xyz, np.float64), _c(mesh.vertices_on_cell, np.int32),
            _c(mesh.cells_on_cell, np.int32), _c(mesh.n_edges_on_cell, np.int32))
    return [C.c_int(mesh.n_cells), C.c_int(mesh.n_vertices), C.c_int(mesh.max_edges), C.c_int(L)] + [_p(a) for a in arrs], arrs


def _depths(n, depth, depths):
    if depths is not None:
        return np.array(depths, dtype=np.float32, copy=True)
    return np.full(n, depth, dtype=np.float32)


def streamline(mesh, prep: Prepared, seeds, cell0, delta_t, duration, record_t, depth=0.0, depths=None,
               method="rk4", direction="forward", log_cells=True):
    lib = _load()
    L = prep.ztop_v.shape[1]
    n = seeds.shape[0]
    each = int(duration) // int(record_t)
    times = int(duration) // int(delta_t)
    pos = np.array(seeds, dtype=np.float64, order="C", copy=True)
    dep = _depths(n, depth, depths)
    out_pos = np.zeros((n, each, 3)); out_vel = np.zeros((n, each, 3))
    cell_log = np.zeros((n, times), dtype=np.int32) if log_cells else None
    steps = np.zeros(n, dtype=np.int32); status = np.zeros(n, dtype=np.int32); fcell = np.zeros(n, dtype=np.int32)
    margs, keep = _mesh_args(mesh, L)
    rc = lib.orc_streamline(*margs, _p(prep.ztop_v), _p(prep.vel_v), _p(prep.w_v),
                            C.c_int(1 if method == "rk4" else 0), C.c_int(1 if direction == "forward" else 0),
                            C.c_int64(delta_t), C.c_int64(duration), C.c_int64(record_t),
                            C.c_int64(n), _p(pos), _p(dep), _p(_c(cell0, np.int32)),
                            _p(out_pos), _p(out_vel), _p(cell_log), _p(steps), _p(status), _p(fcell))
    assert rc == 0, rc
    return {"raw_pos": out_pos, "raw_vel": out_vel, "pos": pos, "depth": dep, "cell_log": cell_log,
            "steps_alive": steps, "status": status, "final_cell": fcell}


def pathline(mesh, front: Prepared, back: Prepared, seeds, cell0, delta_t, duration, record_t, depth=0.0, depths=None,
             method="rk4", direction="forward", log_cells=True):
    lib = _load()
    L = front.ztop_v.shape[1]
    n = seeds.shape[0]
    each = int(duration) // int(record_t)
    n_steps = int(duration) // int(delta_t)
    pos = np.array(seeds, dtype=np.float64, order="C", copy=True)
    dep = _depths(n, depth, depths)
    out_pos = np.zeros((n, each, 3)); out_vel = np.zeros((n, each, 3)); out_attr = np.zeros((n, each, 3))
    cell_log = np.zeros((n, n_steps), dtype=np.int32) if log_cells else None
    steps = np.zeros(n, dtype=np.int32); status = np.zeros(n, dtype=np.int32); fcell = np.zeros(n, dtype=np.int32)
    fa0, fa1, nf = front.attr_list()
    ba0, ba1, nb = back.attr_list()
    # VK:1093-1104: attributes only when the front snapshot holds more than one
    attr_count = min(nf, nb) if nf > 1 else 0
    margs, keep = _mesh_args(mesh, L)
    rc = lib.orc_pathline(*margs,
                          _p(front.ztop_v), _p(front.vel_v), _p(front.w_v), _p(fa0), _p(fa1),
                          _p(back.ztop_v), _p(back.vel_v), _p(back.w_v), _p(ba0), _p(ba1), C.c_int(attr_count),
                          C.c_int(1 if method == "rk4" else 0), C.c_int(1 if direction == "forward" else 0),
                          C.c_int64(delta_t), C.c_int64(duration), C.c_int64(record_t),
                          C.c_int64(n), _p(pos), _p(dep), _p(_c(cell0, np.int32)),
                          _p(out_pos), _p(out_vel), _p(out_attr), _p(cell_log), _p(steps), _p(status), _p(fcell))
    assert rc == 0, rc
    return {"raw_pos": out_pos, "raw_vel": out_vel, "raw_attr": out_attr, "pos": pos, "depth": dep,
            "cell_log": cell_log, "steps_alive": steps, "status": status, "final_cell": fcell}


def remap(mesh, prep: Prepared, width, height, lat_range=(-90.0, 90.0), lon_range=(-180.0, 180.0), depth=800.0,
          pixel_cell=None):
    lib = _load()
    L = prep.ztop_v.shape[1]
    a0, a1, na = prep.attr_list()
    img0 = np.zeros((height, width, 4))
    img1 = np.zeros((height, width, 4)) if na > 1 else None
    have = pixel_cell is not None
    cells = np.array(pixel_cell, dtype=np.int32, copy=True) if have else np.zeros(width * height, dtype=np.int32)
    margs, keep = _mesh_args(mesh, L)
    rc = lib.orc_remap_fixed_depth(*margs, _p(prep.ztop_v), _p(prep.vel_v), _p(a0), _p(a1), C.c_int(na),
                                   C.c_int(width), C.c_int(height), C.c_double(lat_range[0]), C.c_double(lat_range[1]),
                                   C.c_double(lon_range[0]), C.c_double(lon_range[1]), C.c_double(depth),
                                   _p(img0), _p(img1), _p(cells), C.c_int(1 if have else 0))
    assert rc == 0
    return {"img0": img0, "img1": img1, "pixel_cell": cells.reshape(height, width)}


def remap_fixed_layer(mesh, prep: Prepared, width, height, layer, lat_range=(-90.0, 90.0), lon_range=(-180.0, 180.0)):
    lib = _load()
    L = prep.ztop_v.shape[1]
    img = np.zeros((height, width, 4)); cells = np.zeros(width * height, dtype=np.int32)
    margs, keep = _mesh_args(mesh, L)
    rc = lib.orc_remap_fixed_layer(*margs, _p(prep.vel_v), C.c_int(width), C.c_int(height), C.c_double(lat_range[0]),
                                   C.c_double(lat_range[1]), C.c_double(lon_range[0]), C.c_double(lon_range[1]), C.c_int(int(layer)),
                                   _p(img), _p(cells))
    assert rc == 0
    return {"img": img, "pixel_cell": cells.reshape(height, width)}


def regrid_fixed_latitude(mesh, prep: Prepared, width, height, latitude, depth_min, depth_max, lon_range=(-180.0, 180.0)):
    lib = _load()
    L = prep.ztop_v.shape[1]
    img = np.zeros((height, width, 4)); cells = np.zeros(width * height, dtype=np.int32)
    margs, keep = _mesh_args(mesh, L)
    rc = lib.orc_regrid_fixed_latitude(*margs, _p(prep.ztop_v), _p(prep.vel_v), C.c_int(width), C.c_int(height),
                                       C.c_double(lon_range[0]), C.c_double(lon_range[1]), C.c_double(latitude),
                                       C.c_double(depth_min), C.c_double(depth_max), _p(img), _p(cells))
    assert rc == 0
    return {"img": img, "pixel_cell": cells.reshape(height, width)}


def pixel_positions(width, height, lat_range=(-90.0, 90.0), lon_range=(-180.0, 180.0)) -> np.ndarray:
    lib = _load()
    out = np.zeros((height, width, 3))
    tmp = np.zeros(3)
    for i in range(height):
        for j in range(width):
            lib.orc_pixel_position(width, height, lat_range[0], lat_range[1], lon_range[0], lon_range[1], i, j, _p(tmp))
            out[i, j] = tmp
    return out


def pixel_positions_at(width, height, flat_index, lat_range=(-90.0, 90.0), lon_range=(-180.0, 180.0)) -> np.ndarray:
    """sample points of the pixels with the given flat (row-major) indices"""
    lib = _load()
    idx = np.asarray(flat_index, dtype=np.int64)
    out = np.zeros((idx.shape[0], 3))
    tmp = np.zeros(3)
    for k, f in enumerate(idx):
        lib.orc_pixel_position(width, height, lat_range[0], lat_range[1], lon_range[0], lon_range[1], int(f // width), int(f % width), _p(tmp))
        out[k] = tmp
    return out


def finalize_lines(seeds, raw_pos, raw_vel, pathline_mode=False):
    lib = _load()
    n, each = raw_pos.shape[0], raw_pos.shape[1]
    per = each + 1
    pts = np.zeros((n, per, 3)); vel = np.zeros((n, per, 3)); temp = np.zeros((n, per)); sal = np.zeros((n, per))
    last = np.zeros((n, 3))
    lib.orc_finalize_lines(C.c_int64(n), C.c_int(each), _p(_c(seeds, np.float64)), _p(_c(raw_pos, np.float64)),
                           _p(_c(raw_vel, np.float64)), C.c_int(1 if pathline_mode else 0),
                           _p(pts), _p(vel), _p(temp), _p(sal), _p(last))
    return {"points": pts, "velocity": vel, "temperature": temp, "salinity": sal, "last": last}


def wachspress(p, poly) -> np.ndarray:
    lib = _load()
    poly = _c(poly, np.float64)
    w = np.zeros(poly.shape[0])
    lib.orc_wachspress(_p(_c(p, np.float64)), _p(poly), C.c_int(poly.shape[0]), _p(w))
    return w

"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/_ref/libmops_ref.so.

libmops_ref.so is the reference's own TBB/CPU implementation (compiled unmodified from
/root/reference by oracle/build_ref.sh) behind the flat C driver oracle/ref_driver.cpp.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (mops_b200/) never does.

The reference keeps all state in one global `MOPS::app`, so there is one session per
process at a time: `RefOracle(mesh, snapshots)` starts a new one.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import tempfile
from typing import Dict, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libmops_ref.so")

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not available():
        raise RuntimeError(f"{LIB_PATH} not built: run oracle/build_ref.sh where /root/reference exists")
    lib = C.CDLL(LIB_PATH)
    lib.refo_init.argtypes = [C.c_char_p]
    lib.refo_set_mesh.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _f64p, _f64p, _i32p, _i32p, _i32p, _i32p, C.c_void_p]
    lib.refo_add_snapshot.argtypes = [C.c_int, C.c_int, _f64p, _f64p, _f64p, _f64p, C.c_void_p, C.c_int,
                                      C.POINTER(C.c_char_p), C.POINTER(C.c_void_p)]
    lib.refo_activate.argtypes = [C.c_int, C.c_int]
    lib.refo_get_prepared.argtypes = [C.c_int] + [C.c_void_p] * 5
    lib.refo_get_prepared_attr.argtypes = [C.c_int, C.c_char_p, _f64p]
    lib.refo_locate.argtypes = [C.c_int64, _f64p, _i32p]
    lib.refo_generate_seeds.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                        _f64p, C.c_int64]
    lib.refo_generate_seeds.restype = C.c_int64
    lib.refo_streamline.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_int64,
                                    _f64p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    lib.refo_streamline.restype = C.c_int64
    lib.refo_pathline.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_int64,
                                  _f64p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    lib.refo_pathline.restype = C.c_int64
    lib.refo_remap_fixed_depth.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                           C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    lib.refo_remap_fixed_layer.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, _f64p]
    lib.refo_regrid_fixed_latitude.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _f64p]
    lib.refo_gauss3.argtypes = [_f64p, _f64p, _f64p]
    lib.refo_wachspress.argtypes = [_f64p, _f64p, C.c_int, _f64p]
    lib.refo_set_threads.argtypes = [C.c_int]
    _lib = lib
    return lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def max_threads() -> int:
    return int(_load().refo_max_threads())


def set_threads(n: int) -> None:
    _load().refo_set_threads(int(n))


def gauss3(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    x = np.zeros(3)
    _load().refo_gauss3(np.ascontiguousarray(a, dtype=np.float64).reshape(9), np.ascontiguousarray(b, dtype=np.float64), x)
    return x


def wachspress(p: np.ndarray, poly: np.ndarray) -> np.ndarray:
    poly = np.ascontiguousarray(poly, dtype=np.float64)
    w = np.zeros(poly.shape[0])
    _load().refo_wachspress(np.ascontiguousarray(p, dtype=np.float64), poly, poly.shape[0], w)
    return w


def generate_seeds(nx, ny, lat_range, lon_range, depth=0.0) -> np.ndarray:
    lib = _load()
    cap = int(nx) * int(ny) + 4 * (int(nx) + int(ny)) + 16
    out = np.zeros((cap, 3))
    n = lib.refo_generate_seeds(nx, ny, lat_range[0], lat_range[1], lon_range[0], lon_range[1], depth, out, cap)
    assert n <= cap
    return out[:n].copy()


class RefOracle:
    """One session of the reference: mesh + snapshots -> run streamline / pathline / remap."""

    def __init__(self, mesh, snapshots: Sequence, quiet: bool = True):
        self.lib = _load()
        self.mesh = mesh
        self.n_levels = snapshots[0].n_levels
        self._dir = tempfile.mkdtemp(prefix="mops_ref_cache_")
        self.lib.refo_init(self._dir.encode())
        L = self.n_levels
        ref_bottom = np.ascontiguousarray(np.cumsum(np.full(L, 5000.0 / L)))
        self.lib.refo_set_mesh(mesh.n_cells, mesh.n_vertices, mesh.max_edges, L,
                               np.ascontiguousarray(mesh.cell_xyz), np.ascontiguousarray(mesh.vertex_xyz),
                               np.ascontiguousarray(mesh.vertices_on_cell), np.ascontiguousarray(mesh.cells_on_cell),
                               np.ascontiguousarray(mesh.cells_on_vertex), np.ascontiguousarray(mesh.n_edges_on_cell),
                               _ptr(ref_bottom))
        for sid, s in enumerate(snapshots):
            names = sorted(s.attrs.keys())
            arrs = [np.ascontiguousarray(s.attrs[n], dtype=np.float64) for n in names]
            cn = (C.c_char_p * max(1, len(names)))(*[n.encode() for n in names])
            ca = (C.c_void_p * max(1, len(names)))(*[a.ctypes.data for a in arrs])
            rc = self.lib.refo_add_snapshot(sid, 1000 + sid, np.ascontiguousarray(s.zonal), np.ascontiguousarray(s.meridional),
                                            np.ascontiguousarray(s.layer_thickness), np.ascontiguousarray(s.bottom_depth),
                                            _ptr(np.ascontiguousarray(s.vert_vel_top)), len(names), cn, ca)
            assert rc == 0
        self.lib.refo_end()
        self.n_snapshots = len(snapshots)
        self.activate(0, None)

    def close(self):
        if self._dir and os.path.isdir(self._dir):
            shutil.rmtree(self._dir, ignore_errors=True)
        self._dir = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def activate(self, front: int, back: Optional[int]):
        self.lib.refo_activate(front, -1 if back is None else back)

    def prepared(self, sol_id: int) -> Dict[str, np.ndarray]:
        m, L = self.mesh, self.n_levels
        out = {
            "ztop_vertex": np.zeros((m.n_vertices, L)), "vel_vertex": np.zeros((m.n_vertices, L, 3)),
            "vertvel_vertex": np.zeros((m.n_vertices, L + 1)), "ztop_cell": np.zeros((m.n_cells, L)),
            "vel_cell": np.zeros((m.n_cells, L, 3)),
        }
        rc = self.lib.refo_get_prepared(sol_id, _ptr(out["ztop_vertex"]), _ptr(out["vel_vertex"]), _ptr(out["vertvel_vertex"]),
                                        _ptr(out["ztop_cell"]), _ptr(out["vel_cell"]))
        assert rc == 0
        return out

    def prepared_attr(self, sol_id: int, name: str) -> np.ndarray:
        out = np.zeros((self.mesh.n_vertices, self.n_levels))
        rc = self.lib.refo_get_prepared_attr(sol_id, name.encode(), out)
        assert rc == 0, rc
        return out

    def locate(self, xyz: np.ndarray) -> np.ndarray:
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        out = np.zeros(xyz.shape[0], dtype=np.int32)
        self.lib.refo_locate(xyz.shape[0], xyz, out)
        return out

    def _traj_args(self, method, direction, delta_t, duration, record_t, depth, depths, n):
        d = None if depths is None else np.ascontiguousarray(depths, dtype=np.float32)
        return (1 if method == "rk4" else 0, 1 if direction == "forward" else 0, int(delta_t), int(duration),
                int(record_t), float(depth), _ptr(d), n), d

    def streamline(self, seeds, delta_t, duration, record_t, depth=0.0, depths=None, method="rk4", direction="forward"):
        seeds = np.ascontiguousarray(seeds, dtype=np.float64)
        n = seeds.shape[0]
        per = int(duration) // int(record_t) + 1
        pts = np.zeros((n, per, 3)); vel = np.zeros((n, per, 3)); last = np.zeros((n, 3))
        sec = C.c_double(0.0)
        args, keep = self._traj_args(method, direction, delta_t, duration, record_t, depth, depths, n)
        nl = self.lib.refo_streamline(*args, seeds, _ptr(pts), _ptr(vel), _ptr(last), C.byref(sec))
        assert nl == n, (nl, n)
        return {"points": pts, "velocity": vel, "last": last, "seconds": sec.value}

    def pathline(self, seeds, delta_t, duration, record_t, depth=0.0, depths=None, method="rk4", direction="forward"):
        seeds = np.array(seeds, dtype=np.float64, order="C", copy=True)
        n = seeds.shape[0]
        per = int(duration) // int(record_t) + 1
        pts = np.zeros((n, per, 3)); vel = np.zeros((n, per, 3))
        temp = np.zeros((n, per)); sal = np.zeros((n, per))
        sec = C.c_double(0.0)
        args, keep = self._traj_args(method, direction, delta_t, duration, record_t, depth, depths, n)
        nl = self.lib.refo_pathline(*args, seeds, _ptr(pts), _ptr(vel), _ptr(temp), _ptr(sal), C.byref(sec))
        assert nl == n, (nl, n)
        return {"points": pts, "velocity": vel, "temperature": temp, "salinity": sal, "seeds_out": seeds,
                "seconds": sec.value}

    def remap_fixed_layer(self, width, height, layer, lat_range=(-90.0, 90.0), lon_range=(-180.0, 180.0)):
        img = np.zeros((height, width, 4))
        self.lib.refo_remap_fixed_layer(width, height, lat_range[0], lat_range[1], lon_range[0], lon_range[1], int(layer), img)
        return img

    def regrid_fixed_latitude(self, width, height, latitude, lon_range=(-180.0, 180.0)):
        """depth axis = refBottomDepth.front() .. back() of the grid this session was built with"""
        img = np.zeros((height, width, 4))
        self.lib.refo_regrid_fixed_latitude(width, height, lon_range[0], lon_range[1], float(latitude), img)
        return img

    def remap(self, width, height, lat_range=(-90.0, 90.0), lon_range=(-180.0, 180.0), depth=800.0):
        img0 = np.zeros((height, width, 4)); img1 = np.zeros((height, width, 4))
        sec = C.c_double(0.0)
        n = self.lib.refo_remap_fixed_depth(width, height, lat_range[0], lat_range[1], lon_range[0], lon_range[1],
                                            float(depth), _ptr(img0), _ptr(img1), C.byref(sec))
        return {"img0": img0, "img1": img1 if n > 1 else None, "n_images": n, "seconds": sec.value}

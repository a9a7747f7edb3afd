"""ctypes binding of the C ABI (include/mops_b200.h) exported by mops_b200/libmops_b200.so.

This is the harness tests/ and bench.py use to call the product exactly the way a
reference-side binding would: plain pointers and sizes.  numpy arrays are
passed as HOST pointers (MOPS_MEM_HOST); torch CUDA tensors as DEVICE pointers
(MOPS_MEM_DEVICE).  There is no fallback of any kind: if the shared library is missing or
no CUDA device is usable the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmops_b200.so")

MEM_HOST, MEM_DEVICE = 0, 1
METHOD_RK4, METHOD_EULER = 0, 1
DIR_FORWARD, DIR_BACKWARD = 0, 1
MAX_SLOTS = 4
SEM_REFERENCE, SEM_WALK = 0, 1
STATUS_NAMES = {0: "alive", 1: "bad_cell", 2: "not_in_cell", 3: "bad_column", 4: "zero_velocity",
                5: "above_surface", 6: "bad_setup"}

# every symbol include/mops_b200.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "mops_abi_version", "mops_create", "mops_destroy", "mops_last_error", "mops_host_alloc", "mops_host_free",
    "mops_synchronize", "mops_set_stream", "mops_mark", "mops_elapsed_ms", "mops_set_mesh", "mops_set_snapshot", "mops_set_snapshot_async", "mops_snapshot_wait", "mops_side_wait_event",
    "mops_get_prepared", "mops_locate", "mops_streamline", "mops_pathline", "mops_streamline_submit", "mops_pathline_submit",
    "mops_traj_wait", "mops_finalize_lines", "mops_order_key", "mops_shard_bounds", "mops_get_stream", "mops_get_device",
    "mops_dist_unique_id", "mops_dist_create", "mops_dist_destroy", "mops_dist_last_error", "mops_dist_set_stream", "mops_dist_gather_traj",
    "mops_multi_create", "mops_multi_destroy", "mops_multi_last_error", "mops_multi_device_count", "mops_multi_ctx",
    "mops_multi_set_mesh", "mops_multi_set_snapshot", "mops_multi_snapshot_wait", "mops_multi_streamline", "mops_multi_pathline",
    "mops_remap_fixed_depth", "mops_remap_fixed_layer", "mops_regrid_fixed_latitude", "mops_get_info",
]


class TrajCfg(C.Structure):
    _fields_ = [("method", C.c_int32), ("direction", C.c_int32), ("delta_t", C.c_int64), ("duration", C.c_int64),
                ("record_t", C.c_int64), ("mem", C.c_int32), ("sort_particles", C.c_int32), ("count_near_edge", C.c_int32),
                ("semantics", C.c_int32), ("reserved", C.c_int32 * 2)]


class TrajIO(C.Structure):
    _fields_ = [("n", C.c_int64), ("xyz", C.c_void_p), ("depth", C.c_void_p), ("cell0", C.c_void_p),
                ("out_pos", C.c_void_p), ("out_vel", C.c_void_p), ("out_attr", C.c_void_p), ("out_cell_log", C.c_void_p),
                ("out_status", C.c_void_p), ("out_steps", C.c_void_p), ("out_cell", C.c_void_p), ("out_min_edge", C.c_void_p)]


class TrajStats(C.Structure):
    _fields_ = [("particle_steps", C.c_int64), ("alive_at_end", C.c_int64), ("kernel_ms", C.c_double),
                ("locate_ms", C.c_double), ("total_ms", C.c_double), ("launches", C.c_int32), ("reserved", C.c_int32),
                ("near_edge_particles", C.c_int64), ("above_surface_particles", C.c_int64)]


class RemapCfg(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("lat_min", C.c_double), ("lat_max", C.c_double),
                ("lon_min", C.c_double), ("lon_max", C.c_double), ("fixed_depth", C.c_double), ("mem", C.c_int32),
                ("reserved", C.c_int32 * 3)]


class ViewCfg(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("lat_min", C.c_double), ("lat_max", C.c_double),
                ("lon_min", C.c_double), ("lon_max", C.c_double), ("fixed_layer", C.c_int32), ("mem", C.c_int32),
                ("fixed_latitude", C.c_double), ("depth_min", C.c_double), ("depth_max", C.c_double)]


class RemapStats(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("total_ms", C.c_double), ("nan_pixels", C.c_int64), ("launches", C.c_int32),
                ("n_images", C.c_int32)]


class Info(C.Structure):
    _fields_ = [("device", C.c_int32), ("sm_count", C.c_int32), ("cc_major", C.c_int32), ("cc_minor", C.c_int32),
                ("l2_bytes", C.c_int64), ("hbm_bytes", C.c_int64), ("mesh_bytes", C.c_int64),
                ("snapshot_bytes", C.c_int64 * MAX_SLOTS), ("record_width", C.c_int32), ("n_levels", C.c_int32),
                ("total_launches", C.c_int64), ("nonmonotone_cells", C.c_int32 * MAX_SLOTS)]


_lib = None


def load_library():
    """dlopen libmops_b200.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("MOPS_B200_LIB", LIB_PATH)   # developer override: A/B runs of differently built kernels
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with mops_b200/csrc/build.sh "
                           "(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(path)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.mops_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.mops_destroy.argtypes = [vp]
    lib.mops_destroy.restype = None
    lib.mops_last_error.argtypes = [vp]
    lib.mops_last_error.restype = C.c_char_p
    lib.mops_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    lib.mops_host_free.argtypes = [vp]
    lib.mops_synchronize.argtypes = [vp]
    lib.mops_set_stream.argtypes = [vp, vp]
    lib.mops_mark.argtypes = [vp, i32]
    lib.mops_elapsed_ms.argtypes = [vp, i32, i32, C.POINTER(dbl)]
    lib.mops_set_mesh.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    snap_args = [vp, i32, i32, vp, vp, vp, vp, vp, i32, C.POINTER(vp), i32]
    lib.mops_set_snapshot.argtypes = snap_args
    lib.mops_set_snapshot_async.argtypes = snap_args
    lib.mops_snapshot_wait.argtypes = [vp, i32]
    lib.mops_side_wait_event.argtypes = [vp, vp]
    lib.mops_get_prepared.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    lib.mops_locate.argtypes = [vp, i32, i64, vp, vp]
    lib.mops_streamline.argtypes = [vp, C.POINTER(TrajCfg), i32, C.POINTER(TrajIO), C.POINTER(TrajStats)]
    lib.mops_pathline.argtypes = [vp, C.POINTER(TrajCfg), i32, i32, C.POINTER(TrajIO), C.POINTER(TrajStats)]
    lib.mops_streamline_submit.argtypes = [vp, C.POINTER(TrajCfg), i32, C.POINTER(TrajIO), C.POINTER(i64)]
    lib.mops_pathline_submit.argtypes = [vp, C.POINTER(TrajCfg), i32, i32, C.POINTER(TrajIO), C.POINTER(i64)]
    lib.mops_traj_wait.argtypes = [vp, i64, i32, C.POINTER(TrajStats)]
    lib.mops_order_key.argtypes = [vp, i32, i64, vp, vp]
    lib.mops_shard_bounds.argtypes = [i64, i32, i32, C.POINTER(i64), C.POINTER(i64)]
    lib.mops_shard_bounds.restype = None
    lib.mops_get_stream.argtypes = [vp]
    lib.mops_get_stream.restype = vp
    lib.mops_get_device.argtypes = [vp]
    lib.mops_dist_unique_id.argtypes = [vp]
    lib.mops_dist_create.argtypes = [C.POINTER(vp), vp, i32, i32, vp]
    lib.mops_dist_destroy.argtypes = [vp]
    lib.mops_dist_destroy.restype = None
    lib.mops_dist_last_error.argtypes = [vp]
    lib.mops_dist_last_error.restype = C.c_char_p
    lib.mops_dist_set_stream.argtypes = [vp, vp]
    lib.mops_dist_gather_traj.argtypes = [vp, i32, i64, C.POINTER(i64), vp, i32, vp, vp, vp, vp, i64, vp, vp, vp, vp]
    lib.mops_multi_create.argtypes = [C.POINTER(vp), i32, vp]
    lib.mops_multi_destroy.argtypes = [vp]
    lib.mops_multi_destroy.restype = None
    lib.mops_multi_last_error.argtypes = [vp]
    lib.mops_multi_last_error.restype = C.c_char_p
    lib.mops_multi_device_count.argtypes = [vp]
    lib.mops_multi_ctx.argtypes = [vp, i32]
    lib.mops_multi_ctx.restype = vp
    lib.mops_multi_set_mesh.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.mops_multi_set_snapshot.argtypes = snap_args + [i32]
    lib.mops_multi_snapshot_wait.argtypes = [vp, i32]
    lib.mops_multi_streamline.argtypes = [vp, C.POINTER(TrajCfg), i32, C.POINTER(TrajIO), C.POINTER(TrajStats)]
    lib.mops_multi_pathline.argtypes = [vp, C.POINTER(TrajCfg), i32, i32, C.POINTER(TrajIO), C.POINTER(TrajStats)]
    lib.mops_finalize_lines.argtypes = [i64, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp]
    lib.mops_remap_fixed_depth.argtypes = [vp, C.POINTER(RemapCfg), i32, vp, vp, vp, C.POINTER(RemapStats)]
    lib.mops_remap_fixed_layer.argtypes = [vp, C.POINTER(ViewCfg), i32, vp, vp, C.POINTER(RemapStats)]
    lib.mops_regrid_fixed_latitude.argtypes = [vp, C.POINTER(ViewCfg), i32, vp, vp, C.POINTER(RemapStats)]
    lib.mops_get_info.argtypes = [vp, C.POINTER(Info)]
    _lib = lib
    return lib


class MopsError(RuntimeError):
    pass


def _ptr(a):
    """host numpy array / torch CUDA tensor / None -> address"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous()
        return a.data_ptr()
    raise TypeError(type(a))


def _np(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


class Engine:
    """One context = one GPU.  Thin, explicit wrapper: every method is one C-ABI call."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.mops_create(C.byref(h), device)
        if rc != 0:
            raise MopsError(f"mops_create(device={device}) failed with {rc} "
                            "(-3 = no usable CUDA device; this engine has no CPU fallback)")
        self.h = h
        self.mesh = None
        self.levels = {}
        self._keep = []

    def close(self):
        self.dist_close()
        if getattr(self, "h", None) and not getattr(self, "_borrowed", False):
            self.lib.mops_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise MopsError(f"[{rc}] {self.lib.mops_last_error(self.h).decode()}")

    def synchronize(self):
        self._ck(self.lib.mops_synchronize(self.h))

    def set_stream(self, cuda_stream: Optional[int]):
        self._ck(self.lib.mops_set_stream(self.h, cuda_stream))

    def mark(self, idx: int):
        self._ck(self.lib.mops_mark(self.h, idx))

    def elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_double(0.0)
        self._ck(self.lib.mops_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    # ---- mesh / snapshots -------------------------------------------------------------
    def set_mesh(self, mesh):
        arrs = (_np(mesh.cell_xyz, np.float64), _np(mesh.vertex_xyz, np.float64), _np(mesh.vertices_on_cell, np.int32),
                _np(mesh.cells_on_cell, np.int32), _np(mesh.cells_on_vertex, np.int32), _np(mesh.n_edges_on_cell, np.int32))
        self._ck(self.lib.mops_set_mesh(self.h, mesh.n_cells, mesh.n_vertices, mesh.max_edges, *[_ptr(a) for a in arrs]))
        self.mesh = mesh

    def set_snapshot(self, slot: int, snap, async_: bool = False):
        names = sorted(snap.attrs.keys())  # std::map order (R11)
        used = names[:2]
        arrs = [_np(snap.attrs[n], np.float64) for n in used]
        pa = (C.c_void_p * max(1, len(arrs)))(*[a.ctypes.data for a in arrs])
        ins = (_np(snap.zonal, np.float64), _np(snap.meridional, np.float64), _np(snap.layer_thickness, np.float64),
               _np(snap.bottom_depth, np.float64), _np(snap.vert_vel_top, np.float64))
        fn = self.lib.mops_set_snapshot_async if async_ else self.lib.mops_set_snapshot
        self._ck(fn(self.h, slot, snap.n_levels, *[_ptr(a) for a in ins], len(arrs), pa, len(names)))
        if async_:
            self._keep.append((ins, arrs))  # host buffers must outlive the async copy
        self.levels[slot] = snap.n_levels

    def set_snapshot_raw(self, slot, n_levels, zonal, merid, thick, bottom, wtop, async_=False):
        """pointers straight through (e.g. pinned host buffers); no attributes"""
        fn = self.lib.mops_set_snapshot_async if async_ else self.lib.mops_set_snapshot
        pa = (C.c_void_p * 1)()
        self._ck(fn(self.h, slot, n_levels, zonal, merid, thick, bottom, wtop, 0, pa, 0))
        self.levels[slot] = n_levels

    def side_wait_event(self, cuda_event: int):
        """the side stream (snapshot upload + preprocessing) waits for a cudaEvent_t recorded by the caller"""
        self._ck(self.lib.mops_side_wait_event(self.h, C.c_void_p(cuda_event)))

    def snapshot_wait(self, slot):
        self._ck(self.lib.mops_snapshot_wait(self.h, slot))
        self._keep.clear()

    def get_prepared(self, slot: int, attrs: int = 0):
        m, L = self.mesh, self.levels[slot]
        out = {"ztop_vertex": np.zeros((m.n_vertices, L)), "vel_vertex": np.zeros((m.n_vertices, L, 3)),
               "vertvel_vertex": np.zeros((m.n_vertices, L + 1))}
        a0 = np.zeros((m.n_vertices, L)) if attrs >= 1 else None
        a1 = np.zeros((m.n_vertices, L)) if attrs >= 2 else None
        self._ck(self.lib.mops_get_prepared(self.h, slot, _ptr(out["ztop_vertex"]), _ptr(out["vel_vertex"]),
                                            _ptr(out["vertvel_vertex"]), _ptr(a0), _ptr(a1)))
        out["attr0"], out["attr1"] = a0, a1
        return out

    def info(self) -> Info:
        i = Info()
        self._ck(self.lib.mops_get_info(self.h, C.byref(i)))
        return i

    # ---- point location ------------------------------------------------------------------
    def locate(self, xyz):
        if isinstance(xyz, np.ndarray):
            xyz = _np(xyz, np.float64)
            out = np.zeros(xyz.shape[0], dtype=np.int32)
            self._ck(self.lib.mops_locate(self.h, MEM_HOST, xyz.shape[0], _ptr(xyz), _ptr(out)))
            return out
        import torch
        out = torch.empty(xyz.shape[0], dtype=torch.int32, device=xyz.device)
        self._ck(self.lib.mops_locate(self.h, MEM_DEVICE, xyz.shape[0], _ptr(xyz), _ptr(out)))
        return out

    def order_key(self, xyz):
        """rank of each point's cell along the mesh's Morton curve (what the engine sorts particles by); -1 = no cell"""
        if isinstance(xyz, np.ndarray):
            xyz = _np(xyz, np.float64)
            out = np.zeros(xyz.shape[0], dtype=np.int32)
            self._ck(self.lib.mops_order_key(self.h, MEM_HOST, xyz.shape[0], _ptr(xyz), _ptr(out)))
            return out
        import torch
        out = torch.empty(xyz.shape[0], dtype=torch.int32, device=xyz.device)
        self._ck(self.lib.mops_order_key(self.h, MEM_DEVICE, xyz.shape[0], _ptr(xyz), _ptr(out)))
        return out

    # ---- one process per GPU: NCCL communicator inside the library ----------------------------
    def dist_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        rc = self.lib.mops_dist_unique_id(buf)
        if rc != 0:
            raise MopsError(f"mops_dist_unique_id -> {rc} (libnccl.so.2 not loadable?)")
        return buf.raw

    def dist_create(self, rank: int, world: int, unique_id: bytes):
        h = C.c_void_p()
        rc = self.lib.mops_dist_create(C.byref(h), self.h, rank, world, C.create_string_buffer(unique_id, 128))
        if rc != 0:
            raise MopsError(f"mops_dist_create(rank={rank}, world={world}) -> {rc}")
        self.dist = h
        return h

    def dist_gather_traj(self, root, counts, index, each, pos=None, vel=None, xyz=None, depth=None, n_total=0,
                         out_pos=None, out_vel=None, out_xyz=None, out_depth=None):
        """device tensors in, caller-order device tensors out on `root` (see include/mops_b200.h); asynchronous on the stream"""
        cnt = (C.c_int64 * len(counts))(*[int(c) for c in counts])
        n_local = int(index.shape[0])
        rc = self.lib.mops_dist_gather_traj(self.dist, root, n_local, cnt, _ptr(index), each, _ptr(pos), _ptr(vel), _ptr(xyz), _ptr(depth),
                                            n_total, _ptr(out_pos), _ptr(out_vel), _ptr(out_xyz), _ptr(out_depth))
        if rc != 0:
            raise MopsError(f"mops_dist_gather_traj -> {rc}: {self.lib.mops_dist_last_error(self.dist).decode()}")

    def dist_set_stream(self, cuda_stream):
        self._ck(self.lib.mops_dist_set_stream(self.dist, C.c_void_p(cuda_stream) if cuda_stream else None))

    def dist_close(self):
        if getattr(self, "dist", None):
            self.lib.mops_dist_destroy(self.dist)
            self.dist = None

    # ---- trajectories ----------------------------------------------------------------------
    def _traj(self, path, slots, seeds, delta_t, duration, record_t, depth, depths, cell0, method, direction,
              log_cells, sort_particles, want_attr, near_edge=False, walk=False):
        seeds = np.array(seeds, dtype=np.float64, order="C", copy=True)
        n = seeds.shape[0]
        each = int(duration) // int(record_t) if record_t else 0
        times = int(duration) // int(delta_t) if delta_t else 0
        dep = np.array(depths, dtype=np.float32, copy=True) if depths is not None else np.full(n, depth, dtype=np.float32)
        out_pos = np.zeros((n, max(each, 0), 3)); out_vel = np.zeros((n, max(each, 0), 3))
        out_attr = np.zeros((n, max(each, 0), 3)) if want_attr else None
        log = np.zeros((n, max(times, 0)), dtype=np.int32) if log_cells else None
        status = np.zeros(n, dtype=np.int32); steps = np.zeros(n, dtype=np.int32); fcell = np.zeros(n, dtype=np.int32)
        c0 = _np(cell0, np.int32)
        edge = np.zeros(n) if near_edge else None
        cfg = TrajCfg(METHOD_RK4 if method == "rk4" else METHOD_EULER, DIR_FORWARD if direction == "forward" else DIR_BACKWARD,
                      int(delta_t), int(duration), int(record_t), MEM_HOST, 1 if sort_particles else 0,
                      1 if near_edge else 0, SEM_WALK if walk else SEM_REFERENCE)
        io = TrajIO(n, _ptr(seeds), _ptr(dep), _ptr(c0), _ptr(out_pos), _ptr(out_vel), _ptr(out_attr), _ptr(log),
                    _ptr(status), _ptr(steps), _ptr(fcell), _ptr(edge))
        st = TrajStats()
        if path:
            rc = self.lib.mops_pathline(self.h, C.byref(cfg), slots[0], slots[1], C.byref(io), C.byref(st))
        else:
            rc = self.lib.mops_streamline(self.h, C.byref(cfg), slots[0], C.byref(io), C.byref(st))
        self._ck(rc)
        return {"raw_pos": out_pos, "raw_vel": out_vel, "raw_attr": out_attr, "pos": seeds, "depth": dep, "cell_log": log,
                "status": status, "steps_alive": steps, "final_cell": fcell, "min_edge": edge, "stats": st}

    def streamline(self, slot, seeds, delta_t, duration, record_t, depth=0.0, depths=None, cell0=None, method="rk4",
                   direction="forward", log_cells=False, sort_particles=True, near_edge=False, walk=False):
        return self._traj(False, (slot, slot), seeds, delta_t, duration, record_t, depth, depths, cell0, method, direction,
                          log_cells, sort_particles, False, near_edge, walk)

    def pathline(self, front, back, seeds, delta_t, duration, record_t, depth=0.0, depths=None, cell0=None, method="rk4",
                 direction="forward", log_cells=False, sort_particles=True, want_attr=True, near_edge=False, walk=False):
        return self._traj(True, (front, back), seeds, delta_t, duration, record_t, depth, depths, cell0, method, direction,
                          log_cells, sort_particles, want_attr, near_edge, walk)

    def traj_device(self, path, slots, cfg: TrajCfg, io: TrajIO, want_stats=True) -> Optional[TrajStats]:
        """device-resident call: the caller filled io with torch data_ptr()s and cfg.mem = MEM_DEVICE"""
        st = TrajStats()
        sp = C.byref(st) if want_stats else None
        if path:
            rc = self.lib.mops_pathline(self.h, C.byref(cfg), slots[0], slots[1], C.byref(io), sp)
        else:
            rc = self.lib.mops_streamline(self.h, C.byref(cfg), slots[0], C.byref(io), sp)
        self._ck(rc)
        return st if want_stats else None

    def traj_submit(self, path, slots, cfg: TrajCfg, io: TrajIO) -> int:
        """asynchronous HOST-memory call (cfg.mem = MEM_HOST, pinned buffers): returns a ticket for traj_wait"""
        tk = C.c_int64(0)
        if path:
            rc = self.lib.mops_pathline_submit(self.h, C.byref(cfg), slots[0], slots[1], C.byref(io), C.byref(tk))
        else:
            rc = self.lib.mops_streamline_submit(self.h, C.byref(cfg), slots[0], C.byref(io), C.byref(tk))
        self._ck(rc)
        return tk.value

    def traj_wait(self, ticket: int, what: int = 1, want_stats=True) -> Optional[TrajStats]:
        """what = 0: end points (io.xyz / io.depth) have landed; what = 1: every output has (stats filled)"""
        st = TrajStats()
        self._ck(self.lib.mops_traj_wait(self.h, C.c_int64(ticket), what, C.byref(st) if (want_stats and what == 1) else None))
        return st if (want_stats and what == 1) else None

    def finalize_lines(self, seeds, raw_pos, raw_vel, pathline_mode=False):
        n, each = raw_pos.shape[0], raw_pos.shape[1]
        per = each + 1
        pts = np.zeros((n, per, 3)); vel = np.zeros((n, per, 3)); temp = np.zeros((n, per)); sal = np.zeros((n, per))
        last = np.zeros((n, 3))
        rc = self.lib.mops_finalize_lines(n, each, _ptr(_np(seeds, np.float64)), _ptr(_np(raw_pos, np.float64)),
                                          _ptr(_np(raw_vel, np.float64)), 1 if pathline_mode else 0, _ptr(pts), _ptr(vel),
                                          _ptr(temp), _ptr(sal), _ptr(last))
        if rc != 0:
            raise MopsError(f"mops_finalize_lines -> {rc}")
        return {"points": pts, "velocity": vel, "temperature": temp, "salinity": sal, "last": last}

    # ---- remap -------------------------------------------------------------------------------
    def remap(self, slot, width, height, lat_range=(-90.0, 90.0), lon_range=(-180.0, 180.0), depth=800.0,
              want_attr=True, want_cells=True):
        img0 = np.zeros((height, width, 4))
        img1 = np.zeros((height, width, 4)) if want_attr else None
        cells = np.zeros((height, width), dtype=np.int32) if want_cells else None
        cfg = RemapCfg(width, height, lat_range[0], lat_range[1], lon_range[0], lon_range[1], float(depth), MEM_HOST)
        st = RemapStats()
        self._ck(self.lib.mops_remap_fixed_depth(self.h, C.byref(cfg), slot, _ptr(img0), _ptr(img1), _ptr(cells), C.byref(st)))
        return {"img0": img0, "img1": img1 if st.n_images > 1 else None, "pixel_cell": cells, "stats": st}

    def remap_fixed_layer(self, slot, width, height, layer, lat_range=(-90.0, 90.0), lon_range=(-180.0, 180.0)):
        img = np.zeros((height, width, 4)); cells = np.zeros((height, width), dtype=np.int32)
        cfg = ViewCfg(width, height, lat_range[0], lat_range[1], lon_range[0], lon_range[1], int(layer), MEM_HOST, 0.0, 0.0, 0.0)
        st = RemapStats()
        self._ck(self.lib.mops_remap_fixed_layer(self.h, C.byref(cfg), slot, _ptr(img), _ptr(cells), C.byref(st)))
        return {"img": img, "pixel_cell": cells, "stats": st}

    def regrid_fixed_latitude(self, slot, width, height, latitude, depth_min, depth_max, lon_range=(-180.0, 180.0)):
        img = np.zeros((height, width, 4)); cells = np.zeros((height, width), dtype=np.int32)
        cfg = ViewCfg(width, height, 0.0, 0.0, lon_range[0], lon_range[1], 0, MEM_HOST, float(latitude), float(depth_min), float(depth_max))
        st = RemapStats()
        self._ck(self.lib.mops_regrid_fixed_latitude(self.h, C.byref(cfg), slot, _ptr(img), _ptr(cells), C.byref(st)))
        return {"img": img, "pixel_cell": cells, "stats": st}

    def remap_device(self, slot, cfg: RemapCfg, img0, img1=None, cells=None, want_stats=True):
        st = RemapStats()
        self._ck(self.lib.mops_remap_fixed_depth(self.h, C.byref(cfg), slot, _ptr(img0), _ptr(img1), _ptr(cells),
                                                 C.byref(st) if want_stats else None))
        return st


class MultiEngine:
    """One process, N GPUs (mops_multi_*): mesh + snapshots replicated, trajectory calls shard their seeds over the devices
    and come back in caller order.  `.dev0` is an Engine view of device 0's context (views, point location)."""

    def __init__(self, n_devices: int = 0):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.mops_multi_create(C.byref(self.h), n_devices, None)
        if rc != 0:
            raise MopsError(f"mops_multi_create({n_devices}) failed with {rc} (devices / NCCL missing; there is no CPU fallback)")
        self.n = int(self.lib.mops_multi_device_count(self.h))
        self.dev0 = Engine.__new__(Engine)
        self.dev0.lib = self.lib
        self.dev0.h = C.c_void_p(self.lib.mops_multi_ctx(self.h, 0))
        self.dev0._borrowed = True
        self.dev0.levels = {}
        self.dev0._keep = []
        self.dev0.mesh = None
        self.levels = {}

    def _ck(self, rc):
        if rc != 0:
            raise MopsError(f"[{rc}] {self.lib.mops_multi_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.mops_multi_destroy(self.h)
            self.h = None
            self.dev0.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_mesh(self, mesh):
        arrs = (_np(mesh.cell_xyz, np.float64), _np(mesh.vertex_xyz, np.float64), _np(mesh.vertices_on_cell, np.int32),
                _np(mesh.cells_on_cell, np.int32), _np(mesh.cells_on_vertex, np.int32), _np(mesh.n_edges_on_cell, np.int32))
        self._ck(self.lib.mops_multi_set_mesh(self.h, mesh.n_cells, mesh.n_vertices, mesh.max_edges, *[_ptr(a) for a in arrs]))
        self.mesh = mesh
        self.dev0.mesh = mesh

    def set_snapshot(self, slot: int, snap, async_: bool = False):
        names = sorted(snap.attrs.keys())
        arrs = [_np(snap.attrs[n], np.float64) for n in names[:2]]
        pa = (C.c_void_p * max(1, len(arrs)))(*[a.ctypes.data for a in arrs])
        ins = (_np(snap.zonal, np.float64), _np(snap.meridional, np.float64), _np(snap.layer_thickness, np.float64),
               _np(snap.bottom_depth, np.float64), _np(snap.vert_vel_top, np.float64))
        self._ck(self.lib.mops_multi_set_snapshot(self.h, slot, snap.n_levels, *[_ptr(a) for a in ins], len(arrs), pa, len(names),
                                                  1 if async_ else 0))
        if async_:
            self._ck(self.lib.mops_multi_snapshot_wait(self.h, slot))
        self.levels[slot] = snap.n_levels
        self.dev0.levels[slot] = snap.n_levels

    def _traj(self, path, slots, seeds, delta_t, duration, record_t, depth=0.0, depths=None, cell0=None, method="rk4",
              direction="forward", want_attr=False):
        seeds = np.array(seeds, dtype=np.float64, order="C", copy=True)
        n = seeds.shape[0]
        each = int(duration) // int(record_t) if record_t else 0
        dep = np.array(depths, dtype=np.float32, copy=True) if depths is not None else np.full(n, depth, dtype=np.float32)
        out_pos = np.full((n, max(each, 0), 3), 7.0); out_vel = np.full((n, max(each, 0), 3), 7.0)
        out_attr = np.full((n, max(each, 0), 3), 7.0) if want_attr else None
        status = np.zeros(n, dtype=np.int32); steps = np.zeros(n, dtype=np.int32); fcell = np.zeros(n, dtype=np.int32)
        c0 = _np(cell0, np.int32)
        cfg = TrajCfg(METHOD_RK4 if method == "rk4" else METHOD_EULER, DIR_FORWARD if direction == "forward" else DIR_BACKWARD,
                      int(delta_t), int(duration), int(record_t), MEM_HOST, 1, 0, SEM_REFERENCE)
        io = TrajIO(n, _ptr(seeds), _ptr(dep), _ptr(c0), _ptr(out_pos), _ptr(out_vel), _ptr(out_attr), None,
                    _ptr(status), _ptr(steps), _ptr(fcell), None)
        st = TrajStats()
        if path:
            rc = self.lib.mops_multi_pathline(self.h, C.byref(cfg), slots[0], slots[1], C.byref(io), C.byref(st))
        else:
            rc = self.lib.mops_multi_streamline(self.h, C.byref(cfg), slots[0], C.byref(io), C.byref(st))
        self._ck(rc)
        return {"raw_pos": out_pos, "raw_vel": out_vel, "raw_attr": out_attr, "pos": seeds, "depth": dep,
                "status": status, "steps_alive": steps, "final_cell": fcell, "stats": st}

    def streamline(self, slot, seeds, delta_t, duration, record_t, **kw):
        return self._traj(False, (slot, slot), seeds, delta_t, duration, record_t, **kw)

    def pathline(self, front, back, seeds, delta_t, duration, record_t, **kw):
        return self._traj(True, (front, back), seeds, delta_t, duration, record_t, **kw)

"""Synthetic MPAS-format fixtures: icosahedral spherical-Voronoi meshes + analytic fields.

The arrays produced here have exactly the names, shapes, index base and padding of an
MPAS-Ocean file as the reference reads it (reference: src/IO/MPASOReader.cpp:141-154,
215-223):

* ``xCell/yCell/zCell`` -> ``cell_xyz [nCells,3]``; ``xVertex...`` -> ``vertex_xyz [nVertices,3]``
* ``verticesOnCell``, ``cellsOnCell`` ``[nCells,maxEdges]`` int32, 1-based, 0-padded, CCW seen
  from outside (so that ``dot(cross(v_k, v_k+1), p) >= 0`` inside the cell -- the test the
  reference's ``IsInMesh`` applies, src/CPU/TBB/Kernel/TBBKernel.h:21-54)
* ``cellsOnVertex [nVertices,3]`` int32 1-based; ``nEdgesOnCell [nCells]`` int32
* ``layerThickness``, ``velocityZonal``, ``velocityMeridional`` ``[nCells,nVertLevels]``,
  ``bottomDepth [nCells]``, ``vertVelocityTop [nCells,nVertLevels+1]``, ``refBottomDepth``

Level n of the recursive icosahedron bisection has 10*4**n + 2 cells and 20*4**n vertices:
level 6/7/8/9 = 40,962 / 163,842 / 655,362 / 2,621,442 cells (BASELINE.json configs).
Everything is vectorised numpy so level 9 is generated in well under a minute.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

SPHERE_RADIUS = 6371229.0  # MPAS default sphere radius [m]
SEED_RADIUS = 6371010.0    # radius the reference places seeds / pixels on (GeoConverter.hpp:107)


@dataclass
class Mesh:
    n_cells: int
    n_vertices: int
    max_edges: int
    cell_xyz: np.ndarray          # [nCells,3] f64
    vertex_xyz: np.ndarray        # [nVertices,3] f64
    vertices_on_cell: np.ndarray  # [nCells,maxEdges] i32, 1-based, 0 pad
    cells_on_cell: np.ndarray     # [nCells,maxEdges] i32, 1-based, 0 pad
    cells_on_vertex: np.ndarray   # [nVertices,3] i32, 1-based
    n_edges_on_cell: np.ndarray   # [nCells] i32
    level: int = -1

    def cell_latlon(self):
        r = np.linalg.norm(self.cell_xyz, axis=1)
        lat = np.arcsin(self.cell_xyz[:, 2] / r)
        lon = np.arctan2(self.cell_xyz[:, 1], self.cell_xyz[:, 0])
        return lat, lon


@dataclass
class Snapshot:
    zonal: np.ndarray             # [nCells,L]
    meridional: np.ndarray        # [nCells,L]
    layer_thickness: np.ndarray   # [nCells,L]
    bottom_depth: np.ndarray      # [nCells]
    vert_vel_top: np.ndarray      # [nCells,L+1]
    attrs: Dict[str, np.ndarray] = field(default_factory=dict)  # name -> [nCells,L]

    @property
    def n_levels(self) -> int:
        return int(self.zonal.shape[1])


def _icosahedron():
    t = (1.0 + np.sqrt(5.0)) / 2.0
    v = np.array([
        [-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0],
        [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
        [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([
        [0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
        [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
        [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
        [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    return v, f


def _subdivide(v: np.ndarray, f: np.ndarray):
    """One 1->4 bisection of every triangle; midpoints shared between faces."""
    n = v.shape[0]
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
    lo = np.minimum(e[:, 0], e[:, 1])
    hi = np.maximum(e[:, 0], e[:, 1])
    key = lo * np.int64(n) + hi
    uniq, inv = np.unique(key, return_inverse=True)
    ulo = uniq // n
    uhi = uniq % n
    mid = v[ulo] + v[uhi]
    mid /= np.linalg.norm(mid, axis=1, keepdims=True)
    v2 = np.concatenate([v, mid], axis=0)
    nf = f.shape[0]
    m01 = n + inv[0:nf]
    m12 = n + inv[nf:2 * nf]
    m20 = n + inv[2 * nf:3 * nf]
    a, b, c = f[:, 0], f[:, 1], f[:, 2]
    f2 = np.concatenate([
        np.stack([a, m01, m20], axis=1),
        np.stack([b, m12, m01], axis=1),
        np.stack([c, m20, m12], axis=1),
        np.stack([m01, m12, m20], axis=1)], axis=0)
    return v2, f2


def icosahedral_mesh(level: int, radius: float = SPHERE_RADIUS, rotate_x: float = 0.3) -> Mesh:
    """Voronoi dual of the level-`level` bisected icosahedron, as MPAS-format arrays.

    Generators are rotated by `rotate_x` rad about x so that no cell centre sits on the
    z-axis (the reference's ENU->XYZ conversion returns 0 there, GeoConverter.hpp:230-236).
    """
    v, f = _icosahedron()
    for _ in range(level):
        v, f = _subdivide(v, f)
    c, s = np.cos(rotate_x), np.sin(rotate_x)
    rot = np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)
    v = v @ rot.T
    v /= np.linalg.norm(v, axis=1, keepdims=True)

    # orient every Delaunay triangle counter-clockwise seen from outside
    a, b, cc = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    nrm = np.cross(b - a, cc - a)
    flip = np.einsum("ij,ij->i", nrm, a) < 0
    f[flip] = f[flip][:, [0, 2, 1]]
    nrm[flip] *= -1.0
    # Voronoi vertex = circumcentre direction of the (spherical) Delaunay triangle
    vv = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)

    n_cells = v.shape[0]
    n_vert = f.shape[0]

    # (cell, face) incidences; corner i of face (p0,p1,p2): next = p_{i+1}, prev = p_{i+2}
    cell = f.reshape(-1)                                   # [3*nf]
    face = np.repeat(np.arange(n_vert, dtype=np.int64), 3)
    nxt2 = np.roll(f, -2, axis=1).reshape(-1)              # second neighbour CCW around `cell`
    # angle of the Voronoi vertex around the cell centre in a local tangent frame
    cpos = v[cell]
    ref = np.where(np.abs(cpos[:, [2]]) < 0.9, np.array([[0.0, 0.0, 1.0]]), np.array([[1.0, 0.0, 0.0]]))
    e1 = np.cross(ref, cpos)
    e1 /= np.linalg.norm(e1, axis=1, keepdims=True)
    e2 = np.cross(cpos, e1)
    d = vv[face] - cpos
    ang = np.arctan2(np.einsum("ij,ij->i", d, e2), np.einsum("ij,ij->i", d, e1))
    order = np.lexsort((ang, cell))
    cell_s, face_s, nxt2_s = cell[order], face[order], nxt2[order]
    counts = np.bincount(cell_s, minlength=n_cells)
    start = np.concatenate([[0], np.cumsum(counts)[:-1]])
    slot = np.arange(cell_s.shape[0]) - start[cell_s]
    max_edges = int(counts.max())
    voc = np.zeros((n_cells, max_edges), dtype=np.int32)
    coc = np.zeros((n_cells, max_edges), dtype=np.int32)
    voc[cell_s, slot] = face_s + 1
    # neighbour across the Voronoi edge (vertex k, vertex k+1) is the 2nd CCW neighbour of face k
    coc[cell_s, slot] = nxt2_s + 1

    return Mesh(
        n_cells=n_cells, n_vertices=n_vert, max_edges=max_edges,
        cell_xyz=np.ascontiguousarray(v * radius),
        vertex_xyz=np.ascontiguousarray(vv * radius),
        vertices_on_cell=voc, cells_on_cell=coc,
        cells_on_vertex=np.ascontiguousarray((f + 1).astype(np.int32)),
        n_edges_on_cell=counts.astype(np.int32), level=level)


def rotation_axis(tilt: float) -> np.ndarray:
    """Unit rotation axis tilted by `tilt` rad from +z towards +x."""
    return np.array([np.sin(tilt), 0.0, np.cos(tilt)])


def solid_body_velocity_xyz(xyz: np.ndarray, speed: float, tilt: float, radius: Optional[float] = None) -> np.ndarray:
    """omega x r for a rotation with equatorial speed `speed` [m/s] about the tilted axis."""
    r = np.linalg.norm(xyz, axis=-1, keepdims=True) if radius is None else radius
    return np.cross(rotation_axis(tilt)[None, :] * (speed / r), xyz)


def solid_body_snapshot(mesh: Mesh, n_levels: int, speed: float, tilt: float = 0.3,
                        total_depth: float = 5000.0, shear: float = 0.0, bumpy: float = 0.0,
                        w_amp: float = 0.0, with_attrs: bool = False) -> Snapshot:
    """Analytic snapshot: solid-body rotation (zonal/meridional at cell centres).

    shear  : layer k velocity is scaled by (1 - shear*k/L)   (0 = identical in all layers)
    bumpy  : bottomDepth = total_depth*(1 - bumpy*g(lat,lon)), layers scaled with it
    w_amp  : vertVelocityTop amplitude [m/s] (0 = BASELINE configs)
    """
    lat, lon = mesh.cell_latlon()
    vel = solid_body_velocity_xyz(mesh.cell_xyz, speed, tilt)
    east = np.stack([-np.sin(lon), np.cos(lon), np.zeros_like(lon)], axis=1)
    north = np.stack([-np.sin(lat) * np.cos(lon), -np.sin(lat) * np.sin(lon), np.cos(lat)], axis=1)
    zon = np.einsum("ij,ij->i", vel, east)
    mer = np.einsum("ij,ij->i", vel, north)
    k = np.arange(n_levels, dtype=np.float64)
    scale = 1.0 - shear * k / n_levels
    zonal = zon[:, None] * scale[None, :]
    merid = mer[:, None] * scale[None, :]
    g = 0.5 * (1.0 + np.sin(3.0 * lon) * np.cos(2.0 * lat))
    bottom = total_depth * (1.0 - bumpy * g)
    thick = np.repeat((bottom / n_levels)[:, None], n_levels, axis=1)
    kp = np.arange(n_levels + 1, dtype=np.float64)
    wtop = w_amp * np.sin(2.0 * lon)[:, None] * np.cos(lat)[:, None] * (1.0 - kp / n_levels)[None, :]
    attrs: Dict[str, np.ndarray] = {}
    if with_attrs:
        attrs["temperature"] = 2.0 + 25.0 * np.cos(lat)[:, None] * np.exp(-k / (0.3 * n_levels))[None, :]
        attrs["salinity"] = 34.0 + 1.5 * np.sin(lon)[:, None] * (k / n_levels)[None, :]
    return Snapshot(
        zonal=np.ascontiguousarray(zonal), meridional=np.ascontiguousarray(merid),
        layer_thickness=np.ascontiguousarray(thick), bottom_depth=np.ascontiguousarray(bottom),
        vert_vel_top=np.ascontiguousarray(wtop),
        attrs={n: np.ascontiguousarray(a) for n, a in attrs.items()})


def latlon_to_xyz(lat_deg: np.ndarray, lon_deg: np.ndarray, radius: float = SEED_RADIUS) -> np.ndarray:
    lat = np.deg2rad(np.asarray(lat_deg, dtype=np.float64))
    lon = np.deg2rad(np.asarray(lon_deg, dtype=np.float64))
    return np.stack([radius * np.cos(lat) * np.cos(lon), radius * np.cos(lat) * np.sin(lon),
                     radius * np.sin(lat)], axis=-1)


def seed_grid(nx: int, ny: int, lat_range, lon_range, radius: float = SEED_RADIUS) -> np.ndarray:
    """Same seeds as the reference's MPASOVisualizer::GenerateSamplePoint
    (src/Core/MPASOVisualizer.cpp:120-149): float-accumulating `for (i = min; i < max; i += step)`."""
    min_lat, max_lat = lat_range
    min_lon, max_lon = lon_range
    i_step = (max_lat - min_lat) / float(nx - 1)
    j_step = (max_lon - min_lon) / float(ny - 1)
    lats, lons = [], []
    i = float(min_lat)
    while i < max_lat:
        j = float(min_lon)
        while j < max_lon:
            lats.append(i)
            lons.append(j)
            j += j_step
        i += i_step
    la = np.array(lats) * (np.pi / 180.0)
    lo = np.array(lons) * (np.pi / 180.0)
    return np.stack([radius * np.cos(la) * np.cos(lo), radius * np.cos(la) * np.sin(lo), radius * np.sin(la)], axis=1)


def gaussian_seeds(n: int, seed: int, sigma_deg: float = 25.0, lat_max: float = 80.0,
                   radius: float = SEED_RADIUS) -> np.ndarray:
    """lat/lon ~ N(0, sigma) truncated to |lat| <= lat_max."""
    rng = np.random.default_rng(seed)
    lat = np.empty(0)
    lon = np.empty(0)
    while lat.shape[0] < n:
        la = rng.normal(0.0, sigma_deg, size=n)
        lo = rng.normal(0.0, sigma_deg, size=n)
        ok = (np.abs(la) <= lat_max) & (np.abs(lo) <= 180.0)
        lat = np.concatenate([lat, la[ok]])
        lon = np.concatenate([lon, lo[ok]])
    return latlon_to_xyz(lat[:n], lon[:n], radius)


def uniform_sphere_seeds(n: int, seed: int, lat_max: float = 80.0, radius: float = SEED_RADIUS) -> np.ndarray:
    """Uniform on the sphere within |lat| < lat_max."""
    rng = np.random.default_rng(seed)
    zmax = np.sin(np.deg2rad(lat_max))
    z = rng.uniform(-zmax, zmax, size=n)
    lon = rng.uniform(-np.pi, np.pi, size=n)
    rxy = np.sqrt(1.0 - z * z)
    return np.stack([radius * rxy * np.cos(lon), radius * rxy * np.sin(lon), radius * z], axis=1)


def dump_fixture(path: str, mesh: Mesh, snaps) -> None:
    """Flat binary fixture the C++ tutorials read (tutorial/fixture.hpp): 'MOPSFIX1', six int32
    sizes, mesh arrays, then per snapshot zonal / meridional / layerThickness / bottomDepth /
    vertVelocityTop and the named scalar attributes (32-byte name + values)."""
    names = sorted(snaps[0].attrs.keys())
    L = snaps[0].n_levels
    with open(path, "wb") as f:
        f.write(b"MOPSFIX1")
        np.array([mesh.n_cells, mesh.n_vertices, mesh.max_edges, L, len(snaps), len(names)], dtype=np.int32).tofile(f)
        for a, dt in ((mesh.cell_xyz, np.float64), (mesh.vertex_xyz, np.float64), (mesh.vertices_on_cell, np.int32),
                      (mesh.cells_on_cell, np.int32), (mesh.cells_on_vertex, np.int32), (mesh.n_edges_on_cell, np.int32)):
            np.ascontiguousarray(a, dtype=dt).tofile(f)
        for s in snaps:
            for a in (s.zonal, s.meridional, s.layer_thickness, s.bottom_depth, s.vert_vel_top):
                np.ascontiguousarray(a, dtype=np.float64).tofile(f)
            for n in names:
                f.write(n.encode().ljust(32, b"\0"))
                np.ascontiguousarray(s.attrs[n], dtype=np.float64).tofile(f)


def write_mpas_files(directory: str, mesh: Mesh, snaps, dates=None, per_file: int = 1, version: int = 2,
                     alias_names: bool = True) -> str:
    """Write the fixture as MPAS-Ocean style NetCDF-3 files + the YAML stream description the reference's
    reader takes (schema of mpas.yaml / tutorial/test.yaml): `<dir>/mesh.nc`, `<dir>/hist.<date>.nc`
    (Time unlimited, `per_file` snapshots each) and `<dir>/stream.yaml`; returns the YAML path.
    alias_names writes the time-varying variables under their `timeMonthly_avg_*` names so that the
    `possible_names` resolution is exercised; tracers are written as float32 (widened on read)."""
    import os
    from scipy.io import netcdf_file
    os.makedirs(directory, exist_ok=True)
    L = snaps[0].n_levels
    lat = np.arcsin(mesh.vertex_xyz[:, 2] / np.linalg.norm(mesh.vertex_xyz, axis=1))
    lon = np.arctan2(mesh.vertex_xyz[:, 1], mesh.vertex_xyz[:, 0])
    f = netcdf_file(os.path.join(directory, "mesh.nc"), "w", version=version)
    f.createDimension("nCells", mesh.n_cells)
    f.createDimension("nVertices", mesh.n_vertices)
    f.createDimension("maxEdges", mesh.max_edges)
    f.createDimension("vertexDegree", 3)
    f.createDimension("nVertLevels", L)
    for i, n in enumerate("xyz"):
        f.createVariable(f"{n}Cell", "d", ("nCells",))[:] = mesh.cell_xyz[:, i]
        f.createVariable(f"{n}Vertex", "d", ("nVertices",))[:] = mesh.vertex_xyz[:, i]
    f.createVariable("latVertex", "d", ("nVertices",))[:] = lat
    f.createVariable("lonVertex", "d", ("nVertices",))[:] = lon
    f.createVariable("verticesOnCell", "i", ("nCells", "maxEdges"))[:] = mesh.vertices_on_cell
    f.createVariable("cellsOnCell", "i", ("nCells", "maxEdges"))[:] = mesh.cells_on_cell
    f.createVariable("cellsOnVertex", "i", ("nVertices", "vertexDegree"))[:] = mesh.cells_on_vertex
    f.createVariable("nEdgesOnCell", "i", ("nCells",))[:] = mesh.n_edges_on_cell
    f.createVariable("refBottomDepth", "d", ("nVertLevels",))[:] = np.cumsum(np.full(L, 5000.0 / L))
    f.close()

    pre = "timeMonthly_avg_" if alias_names else ""
    dates = dates or [f"0001-{1 + i:02d}-01" for i in range((len(snaps) + per_file - 1) // per_file)]
    names = sorted(snaps[0].attrs.keys())
    for fi, date in enumerate(dates):
        chunk = snaps[fi * per_file:(fi + 1) * per_file]
        f = netcdf_file(os.path.join(directory, f"hist.{date}.nc"), "w", version=version)
        f.createDimension("Time", None)
        f.createDimension("nCells", mesh.n_cells)
        f.createDimension("nVertLevels", L)
        f.createDimension("nVertLevelsP1", L + 1)
        f.createDimension("StrLen", 64)
        xt = f.createVariable("xtime", "c", ("Time", "StrLen"))
        vz = f.createVariable(pre + "velocityZonal", "d", ("Time", "nCells", "nVertLevels"))
        vm = f.createVariable(pre + "velocityMeridional", "d", ("Time", "nCells", "nVertLevels"))
        lt = f.createVariable(pre + "layerThickness", "d", ("Time", "nCells", "nVertLevels"))
        wv = f.createVariable(pre + "vertVelocityTop", "d", ("Time", "nCells", "nVertLevelsP1"))
        f.createVariable("bottomDepth", "d", ("nCells",))[:] = chunk[0].bottom_depth
        tr = {n: f.createVariable(n, "f", ("Time", "nCells", "nVertLevels")) for n in names}
        for t, s in enumerate(chunk):
            stamp = f"{date}_{t:02d}:00:00".ljust(64)
            xt[t] = np.frombuffer(stamp.encode(), dtype="S1")
            vz[t] = s.zonal; vm[t] = s.meridional; lt[t] = s.layer_thickness; wv[t] = s.vert_vel_top
            for n in names:
                tr[n][t] = s.attrs[n].astype(np.float32)
        f.close()

    def var(name, aliases=None, optional=False):
        out = f"        - name: {name}\n"
        if aliases:
            out += "          possible_names:\n" + "".join(f"            - {a}\n" for a in aliases)
        if optional:
            out += "          optional: true\n"
        return out

    yaml = ("stream:\n  name: mpas\n  path_prefix: \"%s\"\n  substreams:\n"
            "    - name: mesh\n      format: netcdf\n      filenames: \"mesh.nc\"\n      static: true\n      vars:\n" % directory)
    for n in ("xCell", "yCell", "zCell", "xVertex", "yVertex", "zVertex", "latVertex", "lonVertex", "nEdgesOnCell", "cellsOnCell",
              "cellsOnVertex", "verticesOnCell", "refBottomDepth"):
        yaml += var(n)
    yaml += "    - name: data\n      format: netcdf\n      filenames: \"hist.*.nc\"   # glob, sorted\n      vars:\n"
    yaml += var("xtime", ["xtime", "xtime_startMonthly"])
    for n in ("velocityMeridional", "velocityZonal", "vertVelocityTop", "layerThickness"):
        yaml += var(n, [n, "timeMonthly_avg_" + n, "timeDaily_avg_" + n], optional=(n == "layerThickness"))
    yaml += var("bottomDepth")
    for n in names:
        yaml += var(n, [n, "timeMonthly_avg_activeTracers_" + n], optional=True)
    path = os.path.join(directory, "stream.yaml")
    with open(path, "w") as fh:
        fh.write(yaml)
    return path


def voronoi_mesh(points_unit: np.ndarray, radius: float = SPHERE_RADIUS) -> Mesh:
    """MPAS-format arrays of the spherical Voronoi diagram of arbitrary generators (scipy's
    SphericalVoronoi): cells with anywhere from 3 to 10+ edges, so maxEdges > 6 and the engine's wider cell
    records (M = 8, M = 20) are exercised.  Same conventions as icosahedral_mesh (CCW vertices, neighbour k
    across the edge (vertex k, vertex k+1), 1-based, 0-padded)."""
    from scipy.spatial import SphericalVoronoi
    pts = np.asarray(points_unit, dtype=np.float64)
    pts = pts / np.linalg.norm(pts, axis=1, keepdims=True)
    sv = SphericalVoronoi(pts, radius=1.0, center=np.zeros(3))
    sv.sort_vertices_of_regions()
    n_cells, n_vert = pts.shape[0], sv.vertices.shape[0]
    counts = np.array([len(r) for r in sv.regions], dtype=np.int32)
    max_edges = int(counts.max())
    voc = np.zeros((n_cells, max_edges), dtype=np.int32)
    coc = np.zeros((n_cells, max_edges), dtype=np.int32)
    edge_cells = {}
    regions = []
    for c, reg in enumerate(sv.regions):
        reg = list(reg)
        v = sv.vertices[reg]
        # make the order counter-clockwise seen from outside
        if np.dot(np.cross(v[0], v[1]), pts[c]) < 0:
            reg = reg[::-1]
        regions.append(reg)
        for k in range(len(reg)):
            e = (min(reg[k], reg[(k + 1) % len(reg)]), max(reg[k], reg[(k + 1) % len(reg)]))
            edge_cells.setdefault(e, []).append(c)
    for c, reg in enumerate(regions):
        voc[c, :len(reg)] = np.array(reg) + 1
        for k in range(len(reg)):
            e = (min(reg[k], reg[(k + 1) % len(reg)]), max(reg[k], reg[(k + 1) % len(reg)]))
            other = [x for x in edge_cells[e] if x != c]
            coc[c, k] = (other[0] + 1) if other else 0
    # cellsOnVertex: the three generators of the Delaunay triangle whose circumcentre the vertex is
    cov = np.zeros((n_vert, 3), dtype=np.int32)
    vcells = [[] for _ in range(n_vert)]
    for c, reg in enumerate(regions):
        for vtx in reg:
            vcells[vtx].append(c)
    for i, lst in enumerate(vcells):
        lst = (lst + lst[:1] * 3)[:3] if lst else [0, 0, 0]
        cov[i] = np.array(lst[:3]) + 1
    return Mesh(n_cells=n_cells, n_vertices=n_vert, max_edges=max_edges, cell_xyz=np.ascontiguousarray(pts * radius),
                vertex_xyz=np.ascontiguousarray(sv.vertices * radius), vertices_on_cell=voc, cells_on_cell=coc,
                cells_on_vertex=cov, n_edges_on_cell=counts, level=-1)


def jittered_icosahedral_points(level: int, jitter: float, seed: int) -> np.ndarray:
    """generators of the icosahedral mesh moved by up to `jitter` x the cell spacing: Voronoi cells with 5..8 edges"""
    m = icosahedral_mesh(level, radius=1.0)
    rng = np.random.default_rng(seed)
    spacing = np.sqrt(4.0 * np.pi / m.n_cells)
    p = m.cell_xyz + rng.normal(scale=jitter * spacing, size=m.cell_xyz.shape)
    return p / np.linalg.norm(p, axis=1, keepdims=True)


def random_sphere_points(n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    p = rng.normal(size=(n, 3))
    return p / np.linalg.norm(p, axis=1, keepdims=True)


def carve_land(mesh: Mesh, keep: np.ndarray) -> Mesh:
    """Ocean-only sub-mesh the way MPAS's cell culler leaves it: cells with keep == False are removed, the
    vertices of the kept cells stay (coast vertices included), ids are renumbered densely, and every
    reference to a removed cell becomes 0 in cellsOnCell / cellsOnVertex (the boundary markers the reference
    tests in MPASOSolutionTBB.cpp:33-40 and skips in MPASOVisualizerKernels.cpp:908-911)."""
    keep = np.asarray(keep, dtype=bool)
    assert keep.shape == (mesh.n_cells,) and keep.any()
    new_cell = np.zeros(mesh.n_cells + 1, dtype=np.int64)        # 1-based old -> 1-based new, 0 = none
    new_cell[1:][keep] = np.arange(1, keep.sum() + 1)
    voc = mesh.vertices_on_cell[keep]
    used = np.zeros(mesh.n_vertices + 1, dtype=bool)
    used[voc.ravel()] = True
    used[0] = False
    new_vert = np.zeros(mesh.n_vertices + 1, dtype=np.int64)
    new_vert[used] = np.arange(1, used.sum() + 1)
    return Mesh(n_cells=int(keep.sum()), n_vertices=int(used.sum()), max_edges=mesh.max_edges,
                cell_xyz=np.ascontiguousarray(mesh.cell_xyz[keep]),
                vertex_xyz=np.ascontiguousarray(mesh.vertex_xyz[used[1:]]),
                vertices_on_cell=new_vert[voc].astype(np.int32),
                cells_on_cell=new_cell[mesh.cells_on_cell[keep]].astype(np.int32),
                cells_on_vertex=new_cell[mesh.cells_on_vertex[used[1:]]].astype(np.int32),
                n_edges_on_cell=mesh.n_edges_on_cell[keep].copy(), level=mesh.level)


def continents_mask(mesh: Mesh, seed: int = 3, n_blobs: int = 7, frac: float = 0.3) -> np.ndarray:
    """keep-mask with irregular 'continents': union of spherical caps and a meridional wall with a gap (so
    the ocean is non-convex and has narrow straits); about `frac` of the cells are land."""
    rng = np.random.default_rng(seed)
    u = mesh.cell_xyz / np.linalg.norm(mesh.cell_xyz, axis=1, keepdims=True)
    land = np.zeros(mesh.n_cells, dtype=bool)
    centres = rng.normal(size=(n_blobs, 3))
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    radii = rng.uniform(0.15, 0.45, n_blobs) * np.sqrt(frac / 0.3)
    for c, r in zip(centres, radii):
        land |= np.arccos(np.clip(u @ c, -1, 1)) < r
    lat = np.arcsin(u[:, 2]); lon = np.arctan2(u[:, 1], u[:, 0])
    land |= (np.abs(lon - 0.5) < 0.06) & (np.abs(lat) < 1.0) & (np.abs(lat - 0.2) > 0.05)   # wall with a strait
    return ~land

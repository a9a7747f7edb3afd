#!/usr/bin/env python
"""pyMOPS on a synthetic MPAS mesh: the call sequence of the reference's Python tutorials
(tutorial/streamLine.py, pathLine.py, reMapping.py via tutorial/pyMOPSAPI.py of YosefQiu/MOPS) with the
grid and solutions fed through the setters instead of netCDF files.

    python tutorial/pyMOPS_demo.py            # needs a CUDA device; tools/pyMOPS/build.sh first
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools", "pyMOPS"))

import pyMOPS  # noqa: E402
from mops_b200 import synthetic as S  # noqa: E402


def make_grid(mesh, n_levels):
    g = pyMOPS.MPASOGrid()
    T = pyMOPS.GridAttributeType
    g.setGridAttribute(T.kCellSize, mesh.n_cells)
    g.setGridAttribute(T.kVertexSize, mesh.n_vertices)
    g.setGridAttribute(T.kMaxEdgesSize, mesh.max_edges)
    g.setGridAttribute(T.kVertLevels, n_levels)
    g.setGridAttribute(T.kVertLevelsP1, n_levels + 1)
    g.setGridAttributesVec3(T.kCellCoord, mesh.cell_xyz)
    g.setGridAttributesVec3(T.kVertexCoord, mesh.vertex_xyz)
    g.setGridAttributesInt(T.kVerticesOnCell, mesh.vertices_on_cell.reshape(-1).astype(np.uint64))
    g.setGridAttributesInt(T.kCellsOnCell, mesh.cells_on_cell.reshape(-1).astype(np.uint64))
    g.setGridAttributesInt(T.kCellsOnVertex, mesh.cells_on_vertex.reshape(-1).astype(np.uint64))
    g.setGridAttributesInt(T.kNumberVertexOnCell, mesh.n_edges_on_cell.astype(np.uint64))
    return g


def make_solution(snap, tag):
    s = pyMOPS.MPASOSolution()
    A, T = pyMOPS.AttributeType, pyMOPS.GridAttributeType
    L = snap.n_levels
    s.setAttribute(T.kVertLevels, L)
    s.setAttribute(T.kVertLevelsP1, L + 1)
    s.setTimestep(tag)
    s.mTimeStamp = f"0001-01-{tag + 1:02d}"
    s.setAttributesDouble(A.kZonalVelocity, snap.zonal.reshape(-1))
    s.setAttributesDouble(A.kMeridionalVelocity, snap.meridional.reshape(-1))
    s.setAttributesDouble(A.kLayerThickness, snap.layer_thickness.reshape(-1))
    s.setAttributesDouble(A.kBottomDepth, snap.bottom_depth)
    s.setVertVelocityTop(snap.vert_vel_top.reshape(-1))
    for name, a in snap.attrs.items():
        s.setDoubleAttribute(name, a.reshape(-1))
    return s


def main():
    mesh = S.icosahedral_mesh(5)
    snaps = [S.solid_body_snapshot(mesh, 20, 0.4 + 0.1 * i, tilt=0.3 + 0.02 * i, with_attrs=True) for i in range(2)]
    grid = make_grid(mesh, 20)
    sols = [make_solution(s, i) for i, s in enumerate(snaps)]
    pyMOPS.MOPS_Init("gpu")
    pyMOPS.MOPS_Begin()
    pyMOPS.MOPS_AddGridMesh(grid)
    for i, s in enumerate(sols):
        pyMOPS.MOPS_AddAttribute(10 + i, s)
    pyMOPS.MOPS_End()

    seeds_cfg = pyMOPS.SeedsSettings()
    seeds_cfg.setSeedsRange((11, 11))
    seeds_cfg.setGeoBox((-60.0, 60.0), (-170.0, 170.0))
    seeds_cfg.setDepth(800.0)
    seeds = pyMOPS.MOPS_GenerateSeedsPoints(seeds_cfg)

    traj = pyMOPS.TrajectorySettings()
    traj.depth = 800.0
    traj.deltaT = 120
    traj.simulationDuration = 6 * 3600
    traj.recordT = 3600
    traj.directionType = pyMOPS.CalcDirection.kForward
    traj.methodType = pyMOPS.CalcMethodType.kRK4

    pyMOPS.MOPS_ActiveAttribute(10)
    lines = pyMOPS.MOPS_RunStreamLine(traj, seeds)
    print("streamline:", len(lines), "lines x", lines[0]["points"].shape)

    pyMOPS.MOPS_ActiveAttribute(10, 11)
    plines = pyMOPS.MOPS_RunPathLine(traj, seeds)
    print("pathline:", len(plines), "lines; last point of line 0", plines[0]["lastPoint"])

    vis = pyMOPS.VisualizationSettings()
    vis.imageSize = (360, 180)
    vis.LatRange = (-90.0, 90.0)
    vis.LonRange = (-180.0, 180.0)
    vis.FixedDepth = 800.0
    vis.VisType = pyMOPS.VisualizeType.kFixedDepth
    pyMOPS.MOPS_ActiveAttribute(10)
    imgs = pyMOPS.MOPS_RunRemapping(vis)
    print("remap:", [im.shape for im in imgs], "speed range", float(np.nanmin(imgs[0][..., 2])), float(np.nanmax(imgs[0][..., 2])))
    pyMOPS.MOPS_PrintTimingSummary()
    return lines, plines, imgs, seeds, mesh, snaps


if __name__ == "__main__":
    main()

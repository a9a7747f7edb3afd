"""GPU: the pyMOPS module (tools/pyMOPS, pybind11 over the C++ drop-in) through the tutorial-style call
sequence, against the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pymops_demo_matches_oracle():
    subprocess.check_call(["bash", os.path.join(ROOT, "tutorial", "build.sh")])
    subprocess.check_call(["bash", os.path.join(ROOT, "tools", "pyMOPS", "build.sh")])
    sys.path.insert(0, os.path.join(ROOT, "tutorial"))
    sys.path.insert(0, os.path.join(ROOT, "tools", "pyMOPS"))
    import pyMOPS_demo
    from mops_b200 import synthetic as S
    from oracle import port_oracle as P
    lines, plines, imgs, seeds, mesh, snaps = pyMOPS_demo.main()
    assert np.array_equal(seeds, S.seed_grid(11, 11, (-60, 60), (-170, 170)))
    p0, p1 = P.prepare(mesh, snaps[0]), P.prepare(mesh, snaps[1])
    cells = P.locate(mesh, seeds)
    b = P.streamline(mesh, p0, seeds, cells, 120, 21600, 3600, depth=800.0, method="rk4")
    f = P.finalize_lines(seeds, b["raw_pos"], b["raw_vel"])
    pts = np.stack([ln["points"] for ln in lines]); vel = np.stack([ln["velocity"] for ln in lines])
    assert [ln["lineID"] for ln in lines] == list(range(len(lines)))
    assert np.linalg.norm(pts - f["points"], axis=2).max() < 1e-6
    assert np.linalg.norm(vel - f["velocity"], axis=2).max() <= 1e-9 * np.linalg.norm(f["velocity"], axis=2).max()
    b = P.pathline(mesh, p0, p1, seeds, cells, 120, 21600, 3600, depth=800.0, method="rk4")
    f = P.finalize_lines(seeds, b["raw_pos"], b["raw_vel"], pathline_mode=True)
    pts = np.stack([ln["points"] for ln in plines])
    assert np.linalg.norm(pts - f["points"], axis=2).max() < 1e-6
    assert np.allclose(np.stack([ln["temperature"] for ln in plines]), f["temperature"], rtol=1e-9, atol=1e-15)
    assert np.linalg.norm(np.stack([ln["lastPoint"] for ln in plines]) - f["last"], axis=1).max() < 1e-6
    ref = P.remap(mesh, p0, 360, 180, depth=800.0)
    assert len(imgs) == 2 and imgs[0].shape == (180, 360, 4)
    assert np.allclose(imgs[0], ref["img0"], rtol=1e-9, atol=1e-12, equal_nan=True)
    assert np.allclose(imgs[1], ref["img1"], rtol=1e-9, atol=1e-9, equal_nan=True)


def test_pymops_regrid_and_fixed_layer():
    sys.path.insert(0, os.path.join(ROOT, "tutorial"))
    sys.path.insert(0, os.path.join(ROOT, "tools", "pyMOPS"))
    import pyMOPS
    import pyMOPS_demo
    from mops_b200 import synthetic as S
    from oracle import port_oracle as P
    mesh = S.icosahedral_mesh(4)
    snap = S.solid_body_snapshot(mesh, 12, 1.5, tilt=0.3, shear=0.3, bumpy=0.2)
    grid = pyMOPS_demo.make_grid(mesh, 12)
    ref_bottom = np.cumsum(np.full(12, 5000.0 / 12))
    grid.setRefBottomDepth(ref_bottom)
    sol = pyMOPS_demo.make_solution(snap, 0)
    pyMOPS.MOPS_Init("gpu")
    pyMOPS.MOPS_Begin()
    pyMOPS.MOPS_AddGridMesh(grid)
    pyMOPS.MOPS_AddAttribute(77, sol)
    pyMOPS.MOPS_End()
    pyMOPS.MOPS_ActiveAttribute(77)
    prep = P.prepare(mesh, snap)
    vis = pyMOPS.VisualizationSettings()
    vis.imageSize = (100, 40)
    vis.LatRange = (-90.0, 90.0)
    vis.LonRange = (-180.0, 180.0)
    vis.FixedLatitude = 25.0
    img = pyMOPS.MOPS_RunReGrid(vis)
    want = P.regrid_fixed_latitude(mesh, prep, 100, 40, 25.0, ref_bottom[0], ref_bottom[-1])
    assert np.allclose(img, want["img"], rtol=1e-9, atol=1e-12, equal_nan=True) and np.isfinite(img[..., 0]).any()
    vis.FixedLayer = 4
    img = pyMOPS.MOPS_RunFixedLayer(vis)
    want = P.remap_fixed_layer(mesh, prep, 100, 40, 4)
    assert np.allclose(img, want["img"], rtol=1e-9, atol=1e-12, equal_nan=True)

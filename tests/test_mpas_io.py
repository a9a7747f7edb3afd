"""CPU: MPAS file ingestion of the drop-in API: synthetic fixtures are written as
NetCDF-3 files + a YAML stream description in the reference's schema, read back through
pyMOPS.MPASOReader (our YAML-subset parser + NetCDF-3 reader) and compared bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
from mops_b200 import synthetic as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pyMOPS():
    lib = os.path.join(ROOT, "mops_b200", "libmops_b200.so")
    if not os.path.exists(lib):
        subprocess.check_call(["bash", os.path.join(ROOT, "mops_b200", "csrc", "build.sh")])
    subprocess.check_call(["bash", os.path.join(ROOT, "tutorial", "build.sh")])
    subprocess.check_call(["bash", os.path.join(ROOT, "tools", "pyMOPS", "build.sh")])
    sys.path.insert(0, os.path.join(ROOT, "tools", "pyMOPS"))
    import pyMOPS as mod
    return mod


@pytest.mark.parametrize("version,per_file", [(1, 1), (2, 2)])
def test_reader_round_trip(pyMOPS, tmp_path, version, per_file):
    m = cases.mesh(3)
    snaps = [S.solid_body_snapshot(m, 7, 0.5 + 0.1 * i, tilt=0.3, shear=0.3, bumpy=0.2, w_amp=1e-3, with_attrs=True) for i in range(4)]
    yaml = S.write_mpas_files(str(tmp_path), m, snaps, per_file=per_file, version=version)
    grid = pyMOPS.MPASOGrid()
    grid.init_from_reader(pyMOPS.MPASOReader.readGridData(yaml))
    assert (grid.mCellsSize, grid.mVertexSize, grid.mMaxEdgesSize, grid.mMeshName) == (m.n_cells, m.n_vertices, m.max_edges, "mesh")
    assert np.array_equal(grid.getArray("cellCoord"), m.cell_xyz)
    assert np.array_equal(grid.getArray("vertexCoord"), m.vertex_xyz)
    assert np.array_equal(grid.getArray("verticesOnCell"), m.vertices_on_cell.reshape(-1))
    assert np.array_equal(grid.getArray("cellsOnCell"), m.cells_on_cell.reshape(-1))
    assert np.array_equal(grid.getArray("cellsOnVertex"), m.cells_on_vertex.reshape(-1))
    assert np.array_equal(grid.getArray("nEdgesOnCell"), m.n_edges_on_cell)
    n_files = (len(snaps) + per_file - 1) // per_file
    for fi in range(n_files):
        date = f"0001-{1 + fi:02d}-01"
        for t in range(per_file):
            k = fi * per_file + t
            sol = pyMOPS.MPASOSolution()
            sol.init_from_reader(pyMOPS.MPASOReader.readSolData(yaml, date, t))
            sol.add_attribute("temperature", pyMOPS.AttributeFormat.kFloat)
            sol.add_attribute("salinity", pyMOPS.AttributeFormat.kFloat)
            assert sol.mVertLevels == 7 and sol.mDataName == f"hist.{date}"
            assert sol.getTimeStamp().strip() == f"{date}_{t:02d}:00:00"
            assert np.array_equal(sol.getArray("velocityZonal"), snaps[k].zonal.reshape(-1))          # via possible_names alias
            assert np.array_equal(sol.getArray("velocityMeridional"), snaps[k].meridional.reshape(-1))
            assert np.array_equal(sol.getArray("layerThickness"), snaps[k].layer_thickness.reshape(-1))
            assert np.array_equal(sol.getArray("vertVelocityTop"), snaps[k].vert_vel_top.reshape(-1))
            assert np.array_equal(sol.getArray("bottomDepth"), snaps[fi * per_file].bottom_depth)
            assert np.array_equal(sol.getArray("temperature"), snaps[k].attrs["temperature"].astype(np.float32).astype(np.float64).reshape(-1))
            assert np.array_equal(sol.getArray("salinity"), snaps[k].attrs["salinity"].astype(np.float32).astype(np.float64).reshape(-1))


def test_hdf5_file_is_rejected_with_a_reason(pyMOPS, tmp_path):
    m = cases.mesh(2)
    snaps = [S.solid_body_snapshot(m, 4, 0.5)]
    yaml = S.write_mpas_files(str(tmp_path), m, snaps)
    with open(os.path.join(str(tmp_path), "mesh.nc"), "r+b") as f:
        f.write(b"\x89HDF\r\n\x1a\n")
    code = ("import sys; sys.path.insert(0, %r); import pyMOPS; pyMOPS.MPASOReader.readGridData(%r)"
            % (os.path.join(ROOT, "tools", "pyMOPS"), yaml))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0 and "netCDF-4/HDF5" in r.stderr


@pytest.mark.parametrize("damage", ["truncated_header", "bad_dimension_id", "huge_name", "truncated_data"])
def test_malformed_netcdf_fails_with_a_message_not_a_crash(pyMOPS, tmp_path, damage):
    """the header parser does not trust the file: lengths are checked against the file size, dimension ids against
    the dimension list, and the reader exits with a message (the reference's reader aborts the process as well)"""
    m = cases.mesh(2)
    snaps = [S.solid_body_snapshot(m, 4, 0.5)]
    yaml = S.write_mpas_files(str(tmp_path), m, snaps)
    path = os.path.join(str(tmp_path), "mesh.nc")
    raw = bytearray(open(path, "rb").read())
    if damage == "truncated_header":
        raw = raw[:40]
    elif damage == "truncated_data":
        raw = raw[: len(raw) // 2]
    elif damage == "huge_name":
        raw[16:20] = (0x7FFFFFF0).to_bytes(4, "big")   # length of the first dimension's name
    else:
        # xCell's first dimension id -> 9999 (name record: length 5, "xCell", 3 pad bytes, then the rank and the ids)
        i = raw.find(b"\x00\x00\x00\x05xCell\x00\x00\x00")
        assert i > 0
        pos = i + 12
        assert int.from_bytes(raw[pos:pos + 4], "big") >= 1
        raw[pos + 4:pos + 8] = (9999).to_bytes(4, "big")
    open(path, "wb").write(bytes(raw))
    code = ("import sys; sys.path.insert(0, %r); import pyMOPS; pyMOPS.MPASOReader.readGridData(%r)"
            % (os.path.join(ROOT, "tools", "pyMOPS"), yaml))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0 and r.returncode > 0, (r.returncode, r.stderr[-500:])   # an exit / exception, not a signal
    assert "netcdf" in r.stderr.lower(), r.stderr[-500:]

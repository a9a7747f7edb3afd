"""CPU: the synthetic MPAS-format fixtures have the properties the reference's kernels rely on."""
import numpy as np
import pytest

import cases
from oracle import port_oracle as P


@pytest.mark.parametrize("level,cells", [(2, 162), (4, 2562), (6, 40962)])
def test_counts_and_orientation(level, cells):
    m = cases.mesh(level)
    assert m.n_cells == cells and m.n_vertices == 2 * (cells - 2) and m.max_edges == 6
    assert (m.n_edges_on_cell == 5).sum() == 12
    c = np.arange(m.n_cells)
    voc, n = m.vertices_on_cell - 1, m.n_edges_on_cell
    for k in range(6):
        valid = k < n
        k1 = np.where(k + 1 < n, k + 1, 0)
        a, b = m.vertex_xyz[voc[c, np.minimum(k, n - 1)]], m.vertex_xyz[voc[c, k1]]
        d = np.einsum("ij,ij->i", np.cross(a, b), m.cell_xyz)
        assert (d[valid] > 0).all()  # CCW seen from outside: IsInMesh(cell centre) is true
    assert (voc[c, 5][n == 5] == -1).all()  # 0-padded rows


def test_voronoi_property():
    """every Voronoi vertex is equidistant from its three cells and no other cell is closer
    (the mesh is the Delaunay dual, so the neighbour walk finds the exact nearest centre)"""
    from scipy.spatial import cKDTree
    m = cases.mesh(5)
    t = cKDTree(m.cell_xyz)
    d, idx = t.query(m.vertex_xyz, k=4)
    assert (np.sort(idx[:, :3], axis=1) == np.sort(m.cells_on_vertex - 1, axis=1)).all()
    assert ((d[:, 2] - d[:, 0]) < 1e-6 * d[:, 0]).all() and (d[:, 3] > d[:, 2] * 1.0001).all()


def test_cells_on_cell_share_the_edge():
    m = cases.mesh(4)
    c = np.arange(m.n_cells)
    voc, coc, n = m.vertices_on_cell - 1, m.cells_on_cell - 1, m.n_edges_on_cell
    for k in range(6):
        valid = k < n
        nb = coc[c, k]
        k1 = np.where(k + 1 < n, k + 1, 0)
        va, vb = voc[c, k], voc[c, k1]
        ok = (voc[nb] == va[:, None]).any(1) & (voc[nb] == vb[:, None]).any(1)
        assert ok[valid].all()
    # symmetric adjacency
    for k in range(6):
        valid = k < n
        nb = coc[c, k]
        assert ((coc[nb] == c[:, None]).any(1))[valid].all()


def test_locate_grid_vs_bruteforce():
    m = cases.mesh(5)
    pts = cases.seeds_random(3000, seed=12)
    assert np.array_equal(P.locate(m, pts), P.locate(m, pts, bruteforce=True))

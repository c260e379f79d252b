"""The Foam-free generators of the reference's non-Cartesian example meshes (sdfibm_b200/meshgen.py): the taylor_couette O-grid
(examples/taylor_couette/system/blockMeshDict) and the hanging-node mesh refineMesh leaves behind
(examples/sedimentation/system/refineMeshDict)."""
import math

import numpy as np

from sdfibm_b200 import meshgen


def _closed(mesh):
    """sum of the outward face-area vectors of every cell (owner: +Sf, neighbour: -Sf)"""
    acc = np.zeros((mesh.n_cells, 3))
    np.add.at(acc, mesh.owner, mesh.Sf)
    np.add.at(acc, mesh.neighbour, -mesh.Sf[: mesh.n_internal])
    return np.abs(acc).max()


def test_ogrid_is_the_five_block_mesh_of_the_blockmeshdict():
    n = 12
    m = meshgen.ogrid_taylor_couette(n)
    assert m.n_cells == 5 * n * n
    # shared block faces merged: (n+1)^2 core points + 4 blocks x n rings x n new points, two layers
    assert m.n_points == 2 * ((n + 1) ** 2 + 4 * n * n)
    assert set(np.diff(m.cp_off)) == {8} and set(np.diff(m.cf_off)) == {6}
    assert m.V.min() > 0 and _closed(m) < 1e-14
    # the outer boundary is the n-gon inscribed in the unit circle: 4 n chords
    assert abs(m.V.sum() - 0.5 * 4 * n * math.sin(2 * math.pi / (4 * n))) < 1e-12
    r = np.hypot(m.points[:, 0], m.points[:, 1])
    assert abs(r.max() - 1.0) < 1e-15
    # upper-triangular face order (polyMesh): internal faces sorted by (owner, neighbour), owner < neighbour
    o, nb = m.owner[: m.n_internal], m.neighbour
    assert (o < nb).all() and (np.diff(o) >= 0).all()
    # unstructured at the four core corners: three cells of three different blocks share each of those edges
    P2 = m.points[: m.n_points // 2, :2]
    for c in ((0.5, 0.0), (0.0, -0.5), (-0.5, 0.0), (0.0, 0.5)):
        k = int(np.argmin(np.hypot(P2[:, 0] - c[0], P2[:, 1] - c[1])))
        cells = {ci for ci in range(m.n_cells) if k in m.cp[m.cp_off[ci]: m.cp_off[ci + 1]]}
        assert len(cells) == 3 and len({ci // (n * n) for ci in cells}) == 3
    # curved cells: the outer blocks are not parallelograms
    assert np.unique(np.round(m.V, 12)).size > n


def test_refined_block_keeps_its_hanging_nodes():
    m = meshgen.refine_2d(8, 6, (0.0, 0.0), (1.0, 0.5), (2.0, 5.0, 0.5, 2.5))
    n_fine = 3 * 4
    assert m.n_cells == 8 * 6 - n_fine + 4 * n_fine
    nv, nf = np.diff(m.cp_off), np.diff(m.cf_off)
    assert sorted(set(nv)) == [8, 10] and sorted(set(nf)) == [6, 7]
    assert (nv == 10).sum() == 2 * (3 + 4) and ((nv == 10) == (nf == 7)).all()   # the coarse cells along the refined patch
    assert m.V.min() > 0 and abs(m.V.sum() - 8 * 6 * 0.5) < 1e-12 and _closed(m) < 1e-14
    assert sorted(np.unique(np.round(m.V, 12))) == [0.125, 0.5]
    # pentagonal front / back faces on the 10-vertex cells
    assert sorted(set(np.diff(m.fp_off))) == [4, 5]

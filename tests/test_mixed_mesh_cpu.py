"""The mixed-cell test mesh (hexahedra + prisms + 7-faced polyhedra) is a valid polyMesh and the oracle handles it."""
import numpy as np

from mixed_mesh import mixed_hex_prism_mesh
from oracle.oracle_py import Oracle
from sdfibm_b200.shapes import make_shape, make_solids


def test_mixed_mesh_is_consistent_and_oracle_volume_is_plausible():
    n = 12
    m = mixed_hex_prism_mesh(n)
    assert abs(m.V.sum() - n ** 3) < 1e-9 and (m.V > 0).all()
    assert sorted(set(np.diff(m.cp_off))) == [6, 8] and sorted(set(np.diff(m.cf_off))) == [5, 6, 7]
    # closed cells: the face area vectors of every cell sum to zero
    acc = np.zeros((m.n_cells, 3))
    np.add.at(acc, m.owner, m.Sf)
    np.subtract.at(acc, m.neighbour, m.Sf[: m.n_internal])
    assert np.abs(acc).max() < 1e-12
    shapes = np.array([make_shape("Sphere", radius=3.2)])
    S = make_solids(1)
    S["pos"] = [(5.3, 6.1, 5.7)]
    r = Oracle(m, False).interact(shapes, S, np.zeros((m.n_cells, 3)), 1e-3, 1.0)
    vol = float((r["As"] * m.V).sum())
    assert abs(vol - 4.0 / 3.0 * np.pi * 3.2 ** 3) < 0.08 * vol      # the apex / pyramid fraction is an approximation

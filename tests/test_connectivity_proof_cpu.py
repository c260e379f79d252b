"""The geometric claim behind `conn_proven` (csrc/sdfibm_cuda.cu, k_solid_prepare): on a complete lattice of identical boxes, a ball
(disc) has a FACE-CONNECTED set of vertex-inside cells — so the reference's flood fill returns all of them and the connectivity
certificate can be skipped.  First test: radii just above the cell diagonal, the ball r + one cell inside the mesh (round 1's
narrower claim); second test: ANY radius (down to a fraction of a cell), balls clipped by the mesh, centres outside it.
Checked against the oracle's real flood fill on anisotropic lattices (aspect 0.3 .. 2), radii from 1e-7 above the bound upwards,
random and exactly vertex- / centre-aligned centres, 2-D and 3-D.  3 000 cases were run when this was written; 300 are kept."""
import numpy as np

from oracle.oracle_py import Oracle, eval_points
from sdfibm_b200.mesh import Mesh
from sdfibm_b200.shapes import make_shape, make_solids


def test_ball_cell_sets_on_box_lattices_are_face_connected():
    bad = n_cases = tight = 0
    for seed in range(300):
        rng = np.random.RandomState(seed)
        two_d = bool(rng.randint(0, 2))
        dx = np.array([1.0, float(rng.uniform(0.3, 2.0)), 1.0 if two_d else float(rng.uniform(0.3, 2.0))]) * float(rng.choice([0.1, 1.0, 0.37]))
        hh = 0.5 * dx
        diag = 2.0 * np.sqrt(hh[0] ** 2 + hh[1] ** 2 + (0.0 if two_d else hh[2] ** 2))
        r = diag * (1.0 + 1e-5) * float(rng.choice([1.0000001, 1.001, rng.uniform(1.0, 1.6)]))
        nd = 2 if two_d else 3
        need = [int(np.ceil(2 * (r + 2.0 * hh[d] * (1.0 + 1e-5)) / dx[d])) + 2 for d in range(nd)]
        n = tuple(need[d] + int(rng.randint(0, 3)) for d in range(nd)) + ((1,) if two_d else ())
        x0 = (float(rng.choice([0.0, -3.0, 17.3])), 0.0, -0.5 * dx[2] if two_d else 0.0)
        mesh = Mesh.hex_block(n, x0=x0, dx=tuple(dx))
        lo, hi = mesh.bounds_min, mesh.bounds_max
        marg = r + 2.0 * hh * (1.0 + 1e-5)
        pos = np.array([rng.uniform(lo[d] + marg[d] * (1 + 1e-9), hi[d] - marg[d] * (1 + 1e-9)) if d < nd else 0.0 for d in range(3)])
        if rng.rand() < 0.4:   # vertex- or centre-aligned along some axes (exact ties)
            for d in range(nd):
                if rng.rand() < 0.6:
                    cand = lo[d] + np.round((pos[d] - lo[d]) / (0.5 * dx[d])) * 0.5 * dx[d]
                    if lo[d] + marg[d] < cand < hi[d] - marg[d]: pos[d] = cand
        shapes = np.array([make_shape("Circle" if two_d else "Sphere", radius=r)])
        S = make_solids(1); S[0]["pos"] = pos
        inside, _ = eval_points(shapes, S[0], mesh.points)
        cp = mesh.cp.reshape(-1, 8)
        members = np.nonzero(inside[cp].any(axis=1))[0]
        res = Oracle(mesh, two_d).interact(shapes, S, np.zeros((mesh.n_cells, 3)), 1.0, 1.0, faithful=True)
        got = np.sort(res["list_cells"])
        n_cases += 1
        tight += r < diag * 1.002
        if not np.array_equal(got, members):
            bad += 1
    assert bad == 0 and n_cases == 300 and tight > 100


def test_any_ball_clipped_or_not_has_a_face_connected_cell_set():
    """The general claim (k_solid_prepare's comment): walk every inside vertex towards the point of the mesh box nearest the centre."""
    bad = n_cases = nonempty = outside = small = 0
    for seed in range(400):
        rng = np.random.RandomState(10_000 + seed)
        two_d = bool(rng.randint(0, 2))
        dx = np.array([1.0, float(rng.uniform(0.3, 2.0)), 1.0 if two_d else float(rng.uniform(0.3, 2.0))]) * float(rng.choice([0.1, 1.0, 0.37]))
        nd = 2 if two_d else 3
        diag = np.sqrt((dx[:nd] ** 2).sum())
        r = diag * float(rng.choice([0.3, 0.6, 1.0, rng.uniform(0.2, 3.0)]))
        n = tuple(int(rng.randint(3, 9)) for _ in range(nd)) + ((1,) if two_d else ())
        x0 = (float(rng.choice([0.0, -3.0, 17.3])), 0.0, -0.5 * dx[2] if two_d else 0.0)
        mesh = Mesh.hex_block(n, x0=x0, dx=tuple(dx))
        lo, hi = mesh.bounds_min, mesh.bounds_max
        pos = np.array([rng.uniform(lo[d] - 0.8 * r, hi[d] + 0.8 * r) if d < nd else 0.0 for d in range(3)])
        if rng.rand() < 0.4:   # exact ties: centres on vertices / cell-centre planes
            for d in range(nd):
                if rng.rand() < 0.6:
                    pos[d] = lo[d] + np.round((pos[d] - lo[d]) / (0.5 * dx[d])) * 0.5 * dx[d]
        shapes = np.array([make_shape("Circle" if two_d else "Sphere", radius=r)])
        S = make_solids(1); S[0]["pos"] = pos
        inside, _ = eval_points(shapes, S[0], mesh.points)
        members = np.nonzero(inside[mesh.cp.reshape(-1, 8)].any(axis=1))[0]
        res = Oracle(mesh, two_d).interact(shapes, S, np.zeros((mesh.n_cells, 3)), 1.0, 1.0, faithful=True)
        n_cases += 1
        nonempty += len(members) > 0
        small += r < diag
        outside += bool(np.any(pos[:nd] < lo[:nd]) or np.any(pos[:nd] > hi[:nd]))
        bad += not np.array_equal(np.sort(res["list_cells"]), members)
    assert bad == 0 and n_cases == 400 and nonempty > 300 and outside > 150 and small > 150


def test_well_resolved_ellipsoids_of_any_orientation_have_a_face_connected_cell_set():
    """The second rule of k_solid_prepare: rho = 1.5 (sqrt 3 / 2) h a_max / a_min^2 with 3 rho^2 < 0.9, body two cells inside the mesh."""
    from sdfibm_b200.shapes import quat_from_euler_xyz_deg
    bad = n_cases = tight = 0
    for seed in range(250):
        rng = np.random.RandomState(20_000 + seed)
        two_d = bool(rng.randint(0, 2))
        nd = 2 if two_d else 3
        dx = np.array([1.0, float(rng.uniform(0.5, 1.0)), 1.0 if two_d else float(rng.uniform(0.5, 1.0))])
        h = float(dx[:nd].max())
        a_min = float(rng.uniform(2.2, 4.0))
        # the largest a_max the rule admits for this a_min: 3 rho^2 < 0.9
        a_cap = np.sqrt(0.3) * a_min * a_min / (1.5 * 0.8660254037844386 * h)
        a_max = float(min(a_cap * rng.choice([0.999, 0.9, 0.7]), 3.0 * a_min))
        if a_max < a_min:
            continue
        semi = [a_max, a_min, float(rng.uniform(a_min, a_max))]
        rng.shuffle(semi)
        need = [int(np.ceil(2 * (a_max + 2.0 * dx[d] * (1 + 1e-5)) / dx[d])) + 2 for d in range(nd)]
        n = tuple(need) + ((1,) if two_d else ())
        mesh = Mesh.hex_block(n, x0=(0.0, 0.0, -0.5 * dx[2] if two_d else 0.0), dx=tuple(dx))
        lo, hi = mesh.bounds_min, mesh.bounds_max
        marg = a_max + 2.0 * dx * (1 + 1e-5)
        pos = np.array([rng.uniform(lo[d] + marg[d] * (1 + 1e-9), hi[d] - marg[d] * (1 + 1e-9)) if d < nd else 0.0 for d in range(3)])
        if two_d:
            shapes = np.array([make_shape("Ellipse", radiusa=semi[0], radiusb=semi[1])])
            euler = (0.0, 0.0, float(rng.uniform(-180, 180)))
        else:
            shapes = np.array([make_shape("Ellipsoid", radiusa=semi[0], radiusb=semi[1], radiusc=semi[2])])
            euler = tuple(rng.uniform(-180, 180, size=3))
        S = make_solids(1); S[0]["pos"] = pos; S[0]["quat"] = quat_from_euler_xyz_deg(euler)
        inside, _ = eval_points(shapes, S[0], mesh.points)
        members = np.nonzero(inside[mesh.cp.reshape(-1, 8)].any(axis=1))[0]
        res = Oracle(mesh, two_d).interact(shapes, S, np.zeros((mesh.n_cells, 3)), 1.0, 1.0, faithful=True)
        n_cases += 1
        tight += a_max > 0.95 * a_cap
        bad += not np.array_equal(np.sort(res["list_cells"]), members)
    assert bad == 0 and n_cases > 150 and tight > 20

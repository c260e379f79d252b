"""Synthetic and example workloads (BASELINE.json configs C1–C5) as plain arrays.

Every case is a dict(mesh, two_d, shapes, solids, U, dt, rhof, name).  Shipped example set-ups are
re-stated from the reference case files (cited per function); C4/C5 follow SURVEY.md §8d.
"""
from __future__ import annotations

import math

import numpy as np

from .mesh import Mesh
from .shapes import make_shape, make_solids, quat_from_euler_xyz_deg


def taylor_green(cc: np.ndarray, L: float) -> np.ndarray:
    """U = (sin kx cos ky, -cos kx sin ky, 0.1 sin kz), k = 2 pi / L  (SURVEY.md §8d)."""
    k = 2.0 * math.pi / L
    U = np.empty_like(cc)
    U[:, 0] = np.sin(k * cc[:, 0]) * np.cos(k * cc[:, 1])
    U[:, 1] = -np.cos(k * cc[:, 0]) * np.sin(k * cc[:, 1])
    U[:, 2] = 0.1 * np.sin(k * cc[:, 2])
    return U


def _case(name, mesh, two_d, shapes, solids, U, dt, rhof=1.0):
    return dict(name=name, mesh=mesh, two_d=two_d, shapes=np.ascontiguousarray(shapes), solids=solids,
                U=np.ascontiguousarray(U), dt=dt, rhof=rhof)


# ---- G1: tool_vof/example (reference tool_vof/example/solidDict:16-148) ----------------------
def g1_solids():
    shapes = np.array([
        make_shape("Circle", radius=0.3),
        make_shape("Circle_Tail", radius=0.3, ratio=1, thickness=0.1),
        make_shape("Ellipse", radiusa=0.3, radiusb=0.2),
        make_shape("Rectangle", radiusa=0.3, radiusb=0.2),
        make_shape("Plane"),
    ])
    spec = [(1, (0.5, 1.5, 0), -45), (1, (0.5, 2.5, 0), 0), (1, (0.5, 3.5, 0), 45),
            (3, (1.5, 1.5, 0), 0), (3, (1.5, 2.5, 0), 30), (3, (1.5, 3.5, 0), 60),
            (2, (2.5, 1.5, 0), 0), (2, (2.5, 2.5, 0), 60), (2, (2.5, 3.5, 0), 120),
            (0, (3.5, 1.0, 0), 0), (0, (3.5, 2.5, 0), 0), (0, (3.5, 3.5, 0), 0),
            (4, (0, 0, 0), 15), (4, (0, 0, 0), -90)]
    S = make_solids(len(spec))
    for i, (si, pos, ez) in enumerate(spec):
        S[i]["shape"] = si
        S[i]["pos"] = pos
        S[i]["quat"] = quat_from_euler_xyz_deg((0, 0, ez))
    return shapes, S


def case_g1(points):
    """14 rotated 2-D solids on the shipped 200x200x1 mesh M1 (points from tests/golden/m1_points.npz)."""
    mesh = Mesh.hex_block_with_points((200, 200, 1), points)
    shapes, S = g1_solids()
    S["vel"] = np.array([0.1, -0.2, 0.0])
    S["omega"] = np.array([0.0, 0.0, 0.5])
    U = taylor_green(mesh.cc, 4.0)
    return _case("g1", mesh, True, shapes, S, U, 1.0)


# ---- C1: flow_past_cylinder central block (reference examples/flow_past_cylinder/re200) -------
def case_c1():
    n = 120
    t = Mesh.hex_block((n, n, 1))
    # blockMesh evaluates x = x0 + (x1 - x0) * i / n; keep that expression (never assume i*dx bit patterns)
    ii = np.arange(n + 1)
    xs = -6.0 + 12.0 * ii / n
    P = t.points.copy()
    ix = (np.arange(t.n_points) % (n + 1))
    iy = (np.arange(t.n_points) // (n + 1)) % (n + 1)
    iz = np.arange(t.n_points) // ((n + 1) * (n + 1))
    P[:, 0] = xs[ix]
    P[:, 1] = xs[iy]
    P[:, 2] = -0.5 + iz
    mesh = Mesh.hex_block_with_points((n, n, 1), P)
    shapes = np.array([make_shape("Circle", radius=1.0)])
    S = make_solids(1)
    U = np.zeros((mesh.n_cells, 3))
    U[:, 0] = 1.0
    return _case("c1", mesh, True, shapes, S, U, 0.01)


# ---- C2: sedimentation (reference examples/sedimentation/solidDict:60-1061) -------------------
def case_c2(with_walls=False):
    mesh = Mesh.hex_block((400, 400, 1), (-4.0, 0.0, -0.5), (0.02, 0.02, 1.0))
    shapes = [make_shape("Circle", radius=0.15), make_shape("Plane")]
    pos = []
    for j in range(10):
        for i in range(10):
            pos.append((-1.8 + 0.4 * i, 4.2 + 0.4 * j, 0.0))
    n = len(pos) + (4 if with_walls else 0)
    S = make_solids(n)
    for i, p in enumerate(pos):
        S[i]["pos"] = p
        S[i]["shape"] = 0
    if with_walls:
        # README.md:52-58 — normal +y into the fluid; left wall euler z=-90, right wall +90
        walls = [((0, 0.0, 0), 0), ((0, 8.0, 0), 180), ((-4.0, 4, 0), -90), ((4.0, 4, 0), 90)]
        for k, (p, ez) in enumerate(walls):
            S[100 + k]["pos"] = p
            S[100 + k]["quat"] = quat_from_euler_xyz_deg((0, 0, ez))
            S[100 + k]["shape"] = 1
    rng = np.random.RandomState(7)
    S["vel"][:100] = 0.1 * rng.standard_normal((100, 3)) * np.array([1, 1, 0])
    S["omega"][:100, 2] = 0.5 * rng.standard_normal(100)
    U = taylor_green(mesh.cc, 8.0)
    return _case("c2" + ("_walls" if with_walls else ""), mesh, True, np.array(shapes), S, U, 5e-4)


# ---- C3: falling ellipse on the shipped mesh M2 (reference examples/falling_ellipse/solidDict) --
def case_c3(points):
    mesh = Mesh.hex_block_with_points((100, 200, 1), points)
    shapes = np.array([make_shape("Ellipse", radiusa=0.3, radiusb=0.15)])
    S = make_solids(1)
    S[0]["pos"] = (0.5, 3.5, 0.0)
    S[0]["quat"] = quat_from_euler_xyz_deg((0, 0, -45))
    S[0]["vel"] = (0.0, -0.3, 0.0)
    S[0]["omega"] = (0.0, 0.0, 1.3)
    U = taylor_green(mesh.cc, 4.0)
    return _case("c3", mesh, True, shapes, S, U, 5e-3)


def case_skewed_2d(n=60, seed=3):
    """Rotated + jittered quad block: non-Cartesian cells standing in for the taylor_couette O-grid
    (reference examples/taylor_couette/system/blockMeshDict:43-59), Circle r = 0.3 with spin."""
    t = Mesh.hex_block((n, n, 1), (-1.0, -1.0, -0.5), (2.0 / n, 2.0 / n, 1.0))
    P = t.points.copy()
    rng = np.random.RandomState(seed)
    h = 2.0 / n
    layer = (n + 1) * (n + 1)
    jit = (rng.rand(layer, 2) - 0.5) * 0.5 * h
    ix = np.arange(layer) % (n + 1)
    iy = np.arange(layer) // (n + 1)
    interior = (ix > 0) & (ix < n) & (iy > 0) & (iy < n)
    jit[~interior] = 0.0
    P[:layer, :2] += jit
    P[layer:, :2] += jit
    th = 0.3
    c, s = math.cos(th), math.sin(th)
    x, y = P[:, 0].copy(), P[:, 1].copy()
    P[:, 0] = c * x - s * y
    P[:, 1] = s * x + c * y
    mesh = Mesh.hex_block_with_points((n, n, 1), P)
    shapes = np.array([make_shape("Circle", radius=0.3), make_shape("Circle_TwoTail", radius=0.2, ratio=1.5, thickness=0.1)])
    S = make_solids(2)
    S[0]["pos"] = (0.1, -0.05, 0.0)
    S[0]["omega"] = (0, 0, 6.28)
    S[1]["pos"] = (-0.45, 0.4, 0.0)
    S[1]["quat"] = quat_from_euler_xyz_deg((0, 0, 20))
    S[1]["shape"] = 1
    S[1]["vel"] = (0.2, 0.1, 0)
    U = taylor_green(mesh.cc, 2.0)
    return _case("skewed2d", mesh, True, shapes, S, U, 1e-3)


def case_taylor_couette(n=50):
    """BASELINE config 3, the reference's own set-up: examples/taylor_couette — the five-block O-grid of its blockMeshDict
    (sdfibm_b200.meshgen.ogrid_taylor_couette: curved, non-orthogonal hexahedra, three blocks meeting at the core corners) with the
    solidDict's one Circle r = 0.3 at the origin (:24-29,57-65), free to spin about z; a second, offset tailed shape crosses the
    block junctions so that the unstructured corners are exercised too."""
    from .meshgen import ogrid_taylor_couette

    mesh = ogrid_taylor_couette(n)
    shapes = np.array([make_shape("Circle", radius=0.3), make_shape("Circle_Tail", radius=0.12, ratio=1.6, thickness=0.05)])
    S = make_solids(2)
    S[0]["pos"] = (0.0, 0.0, 0.0)
    S[0]["omega"] = (0, 0, 6.28)
    S[1]["pos"] = (0.33, -0.36, 0.0)
    S[1]["quat"] = quat_from_euler_xyz_deg((0, 0, 35))
    S[1]["shape"] = 1
    S[1]["vel"] = (-0.2, 0.15, 0)
    S[1]["omega"] = (0, 0, -2.0)
    U = taylor_green(mesh.cc, 2.0)
    return _case("taylor_couette", mesh, True, shapes, S, U, 1e-3)


def case_sedimentation_refined(n=40, n_circles=9, seed=4):
    """BASELINE config 2 on the mesh refineMesh leaves behind (examples/sedimentation/system/refineMeshDict + topoSetDict): a
    one-cell-thick block whose central band is split 2 x 2 in the plane, hanging nodes kept — the coarse cells along the band
    have 10 vertices / 7 faces.  Circles of the example's radius-to-cell ratio sit inside the band, across its edge and outside."""
    from .meshgen import refine_2d

    L = 8.0
    h = L / n
    mesh = refine_2d(n, n, (-4.0, 0.0), (h, h), (-2.0, 2.0, 2.0, 6.0))
    shapes = np.array([make_shape("Circle", radius=0.45)])
    rng = np.random.RandomState(seed)
    S = make_solids(n_circles)
    k = int(math.ceil(math.sqrt(n_circles)))
    for i in range(n_circles):
        S[i]["pos"] = (-3.0 + 6.0 * ((i % k) + 0.5) / k + 0.1 * rng.randn(), 1.0 + 6.0 * ((i // k) + 0.5) / k + 0.1 * rng.randn(), 0.0)
        S[i]["vel"] = (0.05 * rng.randn(), -0.3, 0.0)
        S[i]["omega"] = (0, 0, rng.randn())
    U = taylor_green(mesh.cc, 4.0)
    return _case("sedimentation_refined", mesh, True, shapes, S, U, 2e-3)


# ---- random 3-D packs ---------------------------------------------------------------------------
def random_quaternions(rng, n):
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q


def lattice_pack(L, n_side, n_solids, jitter, rng, lo=(0.0, 0.0, 0.0)):
    """First n_solids sites of a seeded shuffle of an n_side^3 lattice of pitch L/n_side, jittered per axis."""
    pitch = L / n_side
    sites = rng.permutation(n_side ** 3)[:n_solids]
    i, j, k = sites % n_side, (sites // n_side) % n_side, sites // (n_side * n_side)
    pos = (np.stack([i, j, k], axis=1) + 0.5) * pitch + rng.uniform(-jitter, jitter, size=(n_solids, 3))
    return pos + np.asarray(lo)


def case_c4(n=256, n_solids=10000, n_side=22, radius=5.0, jitter=0.8, seed=12345):
    """C4: n^3 unit hex cells, spheres r=5 on a jittered lattice, Taylor–Green-like U.
    Scaled-down variants keep pitch/radius: e.g. n=64 -> n_side=5 (pitch 12.8)."""
    mesh = Mesh.hex_block((n, n, n))
    rng = np.random.RandomState(seed)
    pos = lattice_pack(float(n), n_side, n_solids, jitter, rng)
    shapes = np.array([make_shape("Sphere", radius=radius)])
    S = make_solids(n_solids)
    S["pos"] = pos
    S["vel"] = 0.1 * rng.standard_normal((n_solids, 3))
    S["omega"] = 0.05 * rng.standard_normal((n_solids, 3))
    U = taylor_green(mesh.cc, float(n))
    return _case(f"c4_{n}", mesh, False, shapes, S, U, 1e-3)


def c5_solids(L=512.0, n_solids=100000, n_side=47, jitter=0.4, seed=12345):
    """C5 pack: 50% Sphere r=4.5, 50% Ellipsoid (5, 4.5, 4) with random orientation."""
    rng = np.random.RandomState(seed)
    pos = lattice_pack(L, n_side, n_solids, jitter, rng)
    shapes = np.array([make_shape("Sphere", radius=4.5), make_shape("Ellipsoid", radiusa=5.0, radiusb=4.5, radiusc=4.0)])
    S = make_solids(n_solids)
    S["pos"] = pos
    S["shape"] = (np.arange(n_solids) % 2).astype(np.int32)
    S["quat"] = random_quaternions(rng, n_solids)
    S["quat"][S["shape"] == 0] = (1.0, 0.0, 0.0, 0.0)
    S["vel"] = 0.1 * rng.standard_normal((n_solids, 3))
    S["omega"] = 0.05 * rng.standard_normal((n_solids, 3))
    return shapes, S


def decompose_simple(n_ranks):
    """decomposePar `simple` splits used by the scaling config: (2,1,1) / (2,2,1) / (2,2,2)."""
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[n_ranks]


def case_c5_block(rank, n_ranks, n=512, n_solids=100000, n_side=47):
    """Rank `rank`'s subdomain of the n^3 C5 mesh under the simple (x fastest) block split; all solids
    are replicated on every rank (reference: every rank holds every solid)."""
    px, py, pz = decompose_simple(n_ranks)
    ix, iy, iz = rank % px, (rank // px) % py, rank // (px * py)
    nx, ny, nz = n // px, n // py, n // pz
    lo = (ix * nx * 1.0, iy * ny * 1.0, iz * nz * 1.0)
    mesh = Mesh.hex_block((nx, ny, nz), lo)
    shapes, S = c5_solids(float(n), n_solids, n_side)
    U = taylor_green(mesh.cc, float(n))
    return _case(f"c5_{n}_r{rank}of{n_ranks}", mesh, False, shapes, S, U, 1e-3)


def case_mixed3d(n=24, n_solids=40, seed=5):
    """Small 3-D case covering Sphere / Ellipsoid / Box (+com offsets, overlaps) and one Plane wall."""
    mesh = Mesh.hex_block((n, n, n), (0, 0, 0), (1.0, 1.0, 1.0))
    rng = np.random.RandomState(seed)
    shapes = np.array([
        make_shape("Sphere", radius=3.2),
        make_shape("Ellipsoid", radiusa=4.0, radiusb=2.5, radiusc=1.8),
        make_shape("Box", radiusa=2.6, radiusb=1.7, radiusc=3.1),
        make_shape("Sphere", radius=2.0, com=(0.3, -0.2, 0.1)),
        make_shape("Box", radiusa=1.5, radiusb=1.5, radiusc=1.5, com=(0.0, 0.4, 0.0)),
        make_shape("Plane"),
    ])
    S = make_solids(n_solids + 1)
    S["pos"][:n_solids] = rng.uniform(-1.0, n + 1.0, size=(n_solids, 3))
    S["quat"][:n_solids] = random_quaternions(rng, n_solids)
    S["shape"][:n_solids] = rng.randint(0, 5, size=n_solids)
    S["vel"] = 0.3 * rng.standard_normal((n_solids + 1, 3))
    S["omega"] = 0.2 * rng.standard_normal((n_solids + 1, 3))
    S[n_solids]["pos"] = (0.0, 2.3, 0.0)
    S[n_solids]["quat"] = quat_from_euler_xyz_deg((10, 0, 5))
    S[n_solids]["shape"] = 5
    S[n_solids]["vel"] = 0
    S[n_solids]["omega"] = 0
    U = taylor_green(mesh.cc, float(n))
    return _case("mixed3d", mesh, False, shapes, S, U, 1e-3, rhof=1.3)


def prism_mesh(nx, ny, lo=(0.0, 0.0, -0.5), h=(1.0, 1.0, 1.0)):
    """One layer of triangular prisms (every quad split along its diagonal): a non-hex, general-CSR mesh
    with a uniform vertex count (6) per cell."""
    px, py = nx + 1, ny + 1
    pts = np.zeros((2 * px * py, 3))
    for k in range(2):
        for j in range(py):
            for i in range(px):
                pts[i + px * (j + py * k)] = (lo[0] + i * h[0], lo[1] + j * h[1], lo[2] + k * h[2])
    pid = lambda i, j, k: i + px * (j + py * k)  # noqa: E731
    cid = lambda i, j, t: 2 * (i + nx * j) + t   # noqa: E731  (t=0 lower-right triangle, t=1 upper-left)
    faces, owner, neigh = [], [], []

    def add(f, o, nb):
        faces.append(f), owner.append(o), neigh.append(nb)

    for j in range(ny):
        for i in range(nx):
            a, b, c, d = pid(i, j, 0), pid(i + 1, j, 0), pid(i + 1, j + 1, 0), pid(i, j + 1, 0)
            A, B, Cc, D = pid(i, j, 1), pid(i + 1, j, 1), pid(i + 1, j + 1, 1), pid(i, j + 1, 1)
            t0, t1 = cid(i, j, 0), cid(i, j, 1)
            add([a, A, Cc, c], t0, t1)                      # diagonal a-c: normal from t0 (a,b,c) to t1 (a,c,d)
            if i + 1 < nx:
                add([b, c, Cc, B], t0, cid(i + 1, j, 1))    # x+ side of t0 -> neighbour's t1
            else:
                add([b, c, Cc, B], t0, -1)
            if j + 1 < ny:
                add([c, d, D, Cc], t1, cid(i, j + 1, 0))    # y+ side of t1 -> neighbour's t0
            else:
                add([c, d, D, Cc], t1, -1)
            if i == 0:
                add([d, a, A, D], t1, -1)
            if j == 0:
                add([a, b, B, A], t0, -1)
            add([a, c, b], t0, -1)                          # z- caps (normal -z)
            add([a, d, c], t1, -1)
            add([A, B, Cc], t0, -1)                         # z+ caps
            add([A, Cc, D], t1, -1)
    order = sorted(range(len(faces)), key=lambda f: (neigh[f] < 0, f))  # internal faces first
    faces = [faces[f] for f in order]
    owner = [owner[f] for f in order]
    neigh = [neigh[f] for f in order]
    n_int = sum(1 for x in neigh if x >= 0)
    off = np.zeros(len(faces) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(f) for f in faces])
    flat = np.array([v for f in faces for v in f], dtype=np.int32)
    return Mesh.from_polymesh(pts, off, flat, np.array(owner, dtype=np.int32), np.array(neigh[:n_int], dtype=np.int32))


def case_prism2d(n=40):
    mesh = prism_mesh(n, n, (0.0, 0.0, -0.5), (1.0 / n * 4, 1.0 / n * 4, 1.0))
    shapes = np.array([make_shape("Circle", radius=0.8), make_shape("Rectangle", radiusa=0.7, radiusb=0.35),
                       make_shape("Circle_Tail", radius=0.4, ratio=1.2, thickness=0.12)])
    S = make_solids(3)
    S[0]["pos"] = (1.3, 1.4, 0)
    S[1]["pos"] = (2.8, 2.5, 0)
    S[1]["quat"] = quat_from_euler_xyz_deg((0, 0, 33))
    S[1]["shape"] = 1
    S[2]["pos"] = (1.2, 3.0, 0)
    S[2]["quat"] = quat_from_euler_xyz_deg((0, 0, -20))
    S[2]["shape"] = 2
    S["vel"] = [(0.1, 0.2, 0), (-0.3, 0.1, 0), (0, 0, 0)]
    S["omega"] = [(0, 0, 1.0), (0, 0, -2.0), (0, 0, 0.3)]
    U = taylor_green(mesh.cc, 4.0)
    return _case("prism2d", mesh, True, shapes, S, U, 1e-3)


def case_disconnected():
    """A thin rotated rectangle on a coarse mesh: its vertex-inside cell set is NOT face connected, so the
    reference's flood fill keeps only the seed's component — exercises the exact replay."""
    mesh = Mesh.hex_block((24, 24, 1), (0.0, 0.0, -0.5), (1.0, 1.0, 1.0))
    shapes = np.array([make_shape("Rectangle", radiusa=9.0, radiusb=0.18), make_shape("Circle", radius=3.0),
                       make_shape("Circle_Tail", radius=1.2, ratio=6.0, thickness=0.12)])
    S = make_solids(4)
    S[0]["pos"] = (12.1, 11.7, 0)
    S[0]["quat"] = quat_from_euler_xyz_deg((0, 0, 37))
    S[1]["pos"] = (6.3, 16.2, 0)
    S[1]["shape"] = 1
    S[2]["pos"] = (5.2, 4.4, 0)
    S[2]["quat"] = quat_from_euler_xyz_deg((0, 0, 28))
    S[2]["shape"] = 2
    S[3]["pos"] = (17.3, 6.1, 0)
    S[3]["quat"] = quat_from_euler_xyz_deg((0, 0, 118))
    S[3]["shape"] = 0
    S["vel"] = [(0.1, 0.2, 0), (-0.3, 0.1, 0), (0, 0.2, 0), (0.1, 0, 0)]
    S["omega"] = [(0, 0, 1.0), (0, 0, -2.0), (0, 0, 0.3), (0, 0, 0.7)]
    U = taylor_green(mesh.cc, 24.0)
    return _case("disconnected", mesh, True, shapes, S, U, 1e-3)

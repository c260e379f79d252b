"""Random small cases shared by the CPU fuzz (oracle vs the reference's compiled code) and the GPU fuzz (CUDA path vs oracle):
box-cell hex blocks in 2-D / 3-D with several spacings and origins, every shape type with random, integer and half-integer sizes,
solids at random points, exactly on vertices, on cell-centre planes, partly or wholly outside the mesh, aligned and arbitrary
orientations; and non-box cells: jittered + rotated hex blocks, prism meshes, the mixed hex / prism / polyhedron mesh."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mixed_mesh import mixed_hex_prism_mesh  # noqa: E402
from sdfibm_b200 import cases  # noqa: E402
from sdfibm_b200.mesh import Mesh  # noqa: E402
from sdfibm_b200.shapes import make_shape, make_solids, quat_from_euler_xyz_deg  # noqa: E402

DT, RHOF = 2.5e-3, 1.7


def _finish(name, mesh, two_d, specs, S, rng):
    S["shape"] = np.arange(len(specs))
    shapes = np.array([make_shape(t, **kw) for t, kw in specs])
    U = rng.standard_normal((mesh.n_cells, 3))
    return dict(name=name, mesh=mesh, two_d=two_d, shapes=np.ascontiguousarray(shapes), solids=S, U=np.ascontiguousarray(U), dt=DT,
                rhof=RHOF, specs=specs)


def _rand_shape3(rng, h):
    if rng.rand() < 0.2:   # a thin slab: its vertex-inside cell set is usually NOT face connected (flood-fill component only)
        return "Box", dict(radiusa=float(rng.uniform(0.08, 0.4) * h), radiusb=float(rng.uniform(2, 5) * h), radiusc=float(rng.uniform(0.5, 3) * h))
    t = rng.choice(["Sphere", "Ellipsoid", "Box"])
    r = lambda: float(rng.choice([rng.uniform(0.2, 4.0) * h, rng.randint(1, 4) * h, rng.randint(1, 8) * 0.5 * h]))
    if t == "Sphere": return t, dict(radius=r())
    return t, dict(radiusa=r(), radiusb=r(), radiusc=r())
def _rand_shape2(rng, h):
    if rng.rand() < 0.2:   # a needle
        return "Rectangle", dict(radiusa=float(rng.uniform(2, 6) * h), radiusb=float(rng.uniform(0.08, 0.4) * h))
    t = rng.choice(["Circle", "Ellipse", "Rectangle", "Circle_Tail", "Circle_TwoTail", "Plane"])
    r = lambda: float(rng.choice([rng.uniform(0.2, 5.0) * h, rng.randint(1, 5) * h, rng.randint(1, 8) * 0.5 * h]))
    if t == "Circle": return t, dict(radius=r())
    if t in ("Ellipse", "Rectangle"): return t, dict(radiusa=r(), radiusb=r())
    if t == "Plane": return t, dict()
    return t, dict(radius=r(), ratio=float(rng.uniform(0.5, 3)), thickness=float(rng.uniform(0.1, 1.0) * h))

def box_case(seed, two_d):
    rng = np.random.RandomState(seed)
    if two_d:
        n = (int(rng.randint(6, 24)), int(rng.randint(6, 24)), 1); h = float(rng.choice([0.1, 0.25, 1.0, 0.3]))
        mesh = Mesh.hex_block(n, x0=(float(rng.choice([0.0, -1.0, -n[0]*h/2])), float(rng.choice([0.0, -n[1]*h/2])), -0.5), dx=(h, h * float(rng.choice([1.0, 1.0, 0.7])), 1.0))
    else:
        n = tuple(int(x) for x in rng.randint(5, 12, size=3)); h = float(rng.choice([0.1, 0.25, 1.0, 0.3]))
        mesh = Mesh.hex_block(n, x0=tuple(float(x) for x in rng.choice([0.0, -1.0, 0.37], size=3)), dx=(h, h * float(rng.choice([1.0, 0.8])), h * float(rng.choice([1.0, 1.3]))))
    k = int(rng.randint(1, 5))
    specs = [_rand_shape2(rng, h) if two_d else _rand_shape3(rng, h) for _ in range(k)]
    S = make_solids(k)
    lo, hi = mesh.bounds_min, mesh.bounds_max
    for i in range(k):
        mode = rng.randint(0, 4)
        p = rng.uniform(lo - 0.1 * (hi - lo), hi + 0.1 * (hi - lo))
        if mode == 1:   # on a vertex
            p = lo + np.round((p - lo) / h) * h
        elif mode == 2: # on a cell centre along x
            p[0] = lo[0] + (np.floor((p[0] - lo[0]) / h) + 0.5) * h
        if two_d: p[2] = 0.0
        S[i]["pos"] = p
        e = (0, 0, float(rng.choice([0, 0, 45, 90, rng.uniform(-180, 180)]))) if two_d else tuple(float(x) for x in rng.choice([0, 0, 30, 90, rng.uniform(-180, 180)], size=3))
        S[i]["quat"] = quat_from_euler_xyz_deg(e)
        S[i]["vel"] = rng.standard_normal(3) * 0.3; S[i]["omega"] = rng.standard_normal(3) * 0.2
    return _finish(f"box{2 if two_d else 3}d-{seed}", mesh, two_d, specs, S, rng)


def _jittered_block(rng, n, two_d):
    h = 1.0 / max(n)
    t = Mesh.hex_block(n, (0.0, 0.0, -0.5 if two_d else 0.0), (h, h, 1.0 if two_d else h))
    P = t.points.copy()
    nx, ny, nz = n
    idx = np.arange(len(P))
    ix = idx % (nx + 1); iy = (idx // (nx + 1)) % (ny + 1); iz = idx // ((nx + 1) * (ny + 1))
    interior = (ix > 0) & (ix < nx) & (iy > 0) & (iy < ny)
    if two_d:
        layer = (nx + 1) * (ny + 1)
        jit = (rng.rand(layer, 2) - 0.5) * 0.5 * h
        J = np.zeros((len(P), 3)); J[:layer, :2] = jit; J[layer:, :2] = jit
    else:
        interior &= (iz > 0) & (iz < nz)
        J = (rng.rand(len(P), 3) - 0.5) * 0.4 * h
    J[~interior] = 0.0
    P += J
    th = float(rng.choice([0.0, 0.3]))
    c, s = math.cos(th), math.sin(th)
    x, y = P[:, 0].copy(), P[:, 1].copy()
    P[:, 0] = c * x - s * y; P[:, 1] = s * x + c * y
    return Mesh.hex_block_with_points(n, P), h

def general_case(seed, kind):
    rng = np.random.RandomState(seed)
    two_d = kind in ("skew2d", "prism2d")
    if kind == "skew2d":
        mesh, h = _jittered_block(rng, (int(rng.randint(8, 20)), int(rng.randint(8, 20)), 1), True)
    elif kind == "skew3d":
        mesh, h = _jittered_block(rng, tuple(int(x) for x in rng.randint(5, 10, size=3)), False)
    elif kind == "prism2d":
        n = int(rng.randint(6, 16)); h = 0.25
        mesh = cases.prism_mesh(n, n, (0.0, 0.0, -0.5), (h, h, 1.0))
    else:
        mesh = mixed_hex_prism_mesh(int(rng.randint(6, 10))); h = 1.0
    k = int(rng.randint(1, 4))
    specs = [_rand_shape2(rng, h) if two_d else _rand_shape3(rng, h) for _ in range(k)]
    S = make_solids(k)
    lo, hi = mesh.bounds_min, mesh.bounds_max
    for i in range(k):
        p = rng.uniform(lo - 0.05 * (hi - lo), hi + 0.05 * (hi - lo))
        if two_d: p[2] = 0.0
        S[i]["pos"] = p
        e = (0, 0, float(rng.uniform(-180, 180))) if two_d else tuple(float(x) for x in rng.uniform(-180, 180, size=3))
        S[i]["quat"] = quat_from_euler_xyz_deg(e)
        S[i]["vel"] = rng.standard_normal(3) * 0.3; S[i]["omega"] = rng.standard_normal(3) * 0.2
    return _finish(f"{kind}-{seed}", mesh, two_d, specs, S, rng)

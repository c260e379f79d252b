"""Foam-free generators for the NON-Cartesian meshes of the reference's examples (no OpenFOAM here to run blockMesh / refineMesh):

* `ogrid_taylor_couette` — the five-block O-grid of examples/taylor_couette/system/blockMeshDict:17-59: a square core block whose
  corners lie on the axes, four surrounding blocks whose outer edges are circular arcs of radius 1 (the `arc` edges, :50-59),
  every block n x n x 1 cells (the reference: 50).  The blocks meet along shared faces, so the connectivity is unstructured at
  the four core corners (three blocks around an edge) and the outer cells are curved, non-orthogonal hexahedra.
* `refine_2d` — what examples/sedimentation/system/refineMeshDict does to a cellSet of a one-cell-thick block: every selected
  hexahedron is split 2 x 2 in the plane (tan1, tan2; `useHexTopology yes`), and the unrefined neighbours keep the hanging
  nodes — their side face towards the refined region becomes two faces and their front / back faces become pentagons, i.e.
  cells with 10 vertices / 7 faces next to cells with 8 / 6.

Both return a `Mesh` built from points / faces / owner / neighbour in polyMesh conventions (internal faces first, ordered by
owner then neighbour; face normals from owner to neighbour; `cellPoints()` / `cells()` derived by the library's Foam-free mesh
helper exactly as for every other mesh)."""
from __future__ import annotations

import math

import numpy as np

from .mesh import Mesh


def polymesh_from_faces(points, faces):
    """faces: [vertex loop, cell A, cell B or -1].  Orients every loop from owner (the smaller label) to neighbour / outwards,
    sorts internal faces by (owner, neighbour) and boundary faces by owner — OpenFOAM's upper-triangular order."""
    pts = np.asarray(points, dtype=np.float64)
    n_cells = 1 + max(max(f[1], f[2]) for f in faces)
    acc = np.zeros((n_cells, 3))
    cnt = np.zeros(n_cells)
    for loop, ca, cb in faces:
        for c in (ca, cb):
            if c >= 0:
                acc[c] += pts[loop].sum(axis=0)
                cnt[c] += len(loop)
    cen = acc / cnt[:, None]
    out = []
    for loop, ca, cb in faces:
        loop = list(loop)
        if cb >= 0 and cb < ca:
            ca, cb = cb, ca
        p = pts[loop]
        pn = np.roll(p, -1, axis=0)
        nrm = np.array([(p[:, 1] * pn[:, 2] - p[:, 2] * pn[:, 1]).sum(), (p[:, 2] * pn[:, 0] - p[:, 0] * pn[:, 2]).sum(),
                        (p[:, 0] * pn[:, 1] - p[:, 1] * pn[:, 0]).sum()])        # sum of p_q x p_(q+1): twice the area vector
        ref = (cen[cb] - cen[ca]) if cb >= 0 else (p.mean(axis=0) - cen[ca])
        if np.dot(nrm, ref) < 0:
            loop = loop[::-1]
        out.append((loop, ca, cb))
    internal = sorted([f for f in out if f[2] >= 0], key=lambda f: (f[1], f[2]))
    boundary = sorted([f for f in out if f[2] < 0], key=lambda f: f[1])
    allf = internal + boundary
    fp_off = np.zeros(len(allf) + 1, dtype=np.int32)
    fp_off[1:] = np.cumsum([len(f[0]) for f in allf])
    fp = np.concatenate([np.asarray(f[0], dtype=np.int32) for f in allf])
    owner = np.array([f[1] for f in allf], dtype=np.int32)
    neigh = np.array([f[2] for f in internal], dtype=np.int32)
    return Mesh.from_polymesh(pts, fp_off, fp, owner, neigh)


# the six faces of a hexahedron given as 8 point labels in OpenFOAM's hex order (bottom loop 0-3, top loop 4-7)
_HEX_FACES = ((0, 4, 7, 3), (1, 2, 6, 5), (0, 1, 5, 4), (3, 7, 6, 2), (0, 3, 2, 1), (4, 5, 6, 7))


def faces_of_hexes(hexes):
    """Unique faces of a list of hexahedra (8 point labels each): [loop, cell A, cell B or -1]."""
    seen = {}
    faces = []
    for c, h in enumerate(hexes):
        for f in _HEX_FACES:
            loop = [int(h[k]) for k in f]
            key = tuple(sorted(loop))
            t = seen.get(key)
            if t is None:
                seen[key] = len(faces)
                faces.append([loop, c, -1])
            else:
                assert faces[t][2] < 0, "a face shared by more than two cells"
                faces[t][2] = c
    return faces


def ogrid_taylor_couette(n=50, r_core=0.5, r_outer=1.0, z=(-0.5, 0.5)):
    """examples/taylor_couette/system/blockMeshDict:17-59 (see the module docstring)."""
    # block corner points in the plane: 0-3 the core square (on the axes at r_core), 4-7 the outer circle (on the axes at r_outer)
    ang = [-0.5 * math.pi * k for k in range(5)]           # the blocks run clockwise: 0, -90, -180, -270, -360 degrees
    core = [np.array([r_core * round(math.cos(a)), r_core * round(math.sin(a))]) for a in ang[:4]]
    ids = {}
    pts2 = []

    def pid(p):
        key = (round(float(p[0]), 12), round(float(p[1]), 12))
        k = ids.get(key)
        if k is None:
            k = ids[key] = len(pts2)
            pts2.append((float(p[0]), float(p[1])))
        return k

    quads = []   # in-plane quads (counter-clockwise seen from +z), 4 point labels each
    # core block: bilinear between its four corners (blockMeshDict:41, `hex (8 9 10 11 0 1 2 3)`)
    g = np.empty((n + 1, n + 1), dtype=np.int64)
    for j in range(n + 1):
        for i in range(n + 1):
            u, v = i / n, j / n
            p = (1 - u) * (1 - v) * core[0] + u * (1 - v) * core[1] + u * v * core[2] + (1 - u) * v * core[3]
            g[i, j] = pid(p)
    for j in range(n):
        for i in range(n):
            quads.append((g[i, j], g[i + 1, j], g[i + 1, j + 1], g[i, j + 1]))
    # four outer blocks: from core edge (k, k+1) out to the arc between the same two angles (:42-45 with the arcs of :50-59):
    # linear in the radial direction between the core-edge point and the arc point of the same parameter
    for k in range(4):
        a0, a1 = ang[k], ang[k + 1]
        g = np.empty((n + 1, n + 1), dtype=np.int64)
        for i in range(n + 1):
            u = i / n
            inner = (1 - u) * core[k] + u * core[(k + 1) % 4]
            a = (1 - u) * a0 + u * a1
            outer = np.array([r_outer * math.cos(a), r_outer * math.sin(a)])
            outer[np.abs(outer) < 1e-15] = 0.0
            for j in range(n + 1):
                v = j / n
                g[i, j] = pid((1 - v) * inner + v * outer)
        for j in range(n):
            for i in range(n):
                quads.append((g[i, j], g[i, j + 1], g[i + 1, j + 1], g[i + 1, j]))
    P2 = np.array(pts2)
    npl = len(P2)
    points = np.concatenate([np.column_stack([P2, np.full(npl, z[0])]), np.column_stack([P2, np.full(npl, z[1])])])
    hexes = []
    for q in quads:
        # make every quad counter-clockwise so that the extruded hexahedra are right-handed
        x = P2[list(q)]
        area = 0.5 * sum(x[i][0] * x[(i + 1) % 4][1] - x[(i + 1) % 4][0] * x[i][1] for i in range(4))
        q = q if area > 0 else q[::-1]
        hexes.append(list(q) + [p + npl for p in q])
    return polymesh_from_faces(points, faces_of_hexes(hexes))


def refine_2d(nx, ny, x0, dx, region, z=(-0.5, 0.5)):
    """An nx x ny x 1 block (origin x0, spacing dx, both in the plane) whose cells with centres inside `region` =
    (xmin, xmax, ymin, ymax) are split 2 x 2 in the plane, hanging nodes kept (examples/sedimentation/system/refineMeshDict)."""
    ids = {}
    pts2 = []

    def pid(i2, j2):          # lattice of HALF spacings: (i2, j2) = (2 i, 2 j) are the coarse vertices
        k = ids.get((i2, j2))
        if k is None:
            k = ids[(i2, j2)] = len(pts2)
            pts2.append((x0[0] + 0.5 * dx[0] * i2, x0[1] + 0.5 * dx[1] * j2))
        return k

    def fine(i, j):
        cx, cy = x0[0] + dx[0] * (i + 0.5), x0[1] + dx[1] * (j + 0.5)
        return 0 <= i < nx and 0 <= j < ny and region[0] < cx < region[1] and region[2] < cy < region[3]

    # in-plane polygons (counter-clockwise), hanging nodes inserted on the coarse side of a coarse / fine edge
    polys = []
    for j in range(ny):
        for i in range(nx):
            I, J = 2 * i, 2 * j
            if fine(i, j):
                for (a, b) in ((0, 0), (1, 0), (0, 1), (1, 1)):
                    polys.append([pid(I + a, J + b), pid(I + a + 1, J + b), pid(I + a + 1, J + b + 1), pid(I + a, J + b + 1)])
            else:
                loop = [pid(I, J)]
                if fine(i, j - 1):
                    loop.append(pid(I + 1, J))
                loop.append(pid(I + 2, J))
                if fine(i + 1, j):
                    loop.append(pid(I + 2, J + 1))
                loop.append(pid(I + 2, J + 2))
                if fine(i, j + 1):
                    loop.append(pid(I + 1, J + 2))
                loop.append(pid(I, J + 2))
                if fine(i - 1, j):
                    loop.append(pid(I, J + 1))
                polys.append(loop)
    P2 = np.array(pts2)
    npl = len(P2)
    points = np.concatenate([np.column_stack([P2, np.full(npl, z[0])]), np.column_stack([P2, np.full(npl, z[1])])])
    faces = []
    edge_face = {}
    for c, loop in enumerate(polys):
        faces.append([list(loop)[::-1], c, -1])                      # back (z-)
        faces.append([[p + npl for p in loop], c, -1])               # front (z+)
        for k in range(len(loop)):
            a, b = loop[k], loop[(k + 1) % len(loop)]
            key = (min(a, b), max(a, b))
            t = edge_face.get(key)
            if t is None:
                edge_face[key] = len(faces)
                faces.append([[a, b, b + npl, a + npl], c, -1])
            else:
                faces[t][2] = c
    return polymesh_from_faces(points, faces)

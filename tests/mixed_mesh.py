"""Test helper: a conforming mesh that mixes hexahedra, triangular prisms and 7-faced polyhedra."""
import numpy as np

from sdfibm_b200.mesh import Mesh


def mixed_hex_prism_mesh(n=12):
    """An n^3 unit hex block in which every third cell (per axis) is cut into two triangular prisms: the cut cells are prisms,
    their neighbours above / below are 7-faced polyhedra, the rest stay hexahedra."""
    base = Mesh.hex_block((n, n, n))
    pts = base.points.copy()
    n1 = n + 1
    pid = lambda i, j, k: i + n1 * (j + n1 * k)
    cid = lambda i, j, k: i + n * (j + n * k)
    faces = []   # [loop, cellA, cellB]
    for f in range(base.n_faces):
        loop = list(base.fp[base.fp_off[f]:base.fp_off[f + 1]])
        faces.append([loop, int(base.owner[f]), int(base.neighbour[f]) if f < base.n_internal else -1])
    split = [(i, j, k) for k in range(1, n - 1, 3) for j in range(1, n - 1, 3) for i in range(1, n - 1, 3)]
    n_cells = base.n_cells
    key = lambda loop: tuple(sorted(loop))
    index = {key(f[0]): t for t, f in enumerate(faces)}
    dead = set()
    for (i, j, k) in split:
        a, b = cid(i, j, k), n_cells
        n_cells += 1
        P = {(di, dj, dk): pid(i + di, j + dj, k + dk) for di in (0, 1) for dj in (0, 1) for dk in (0, 1)}
        for dk in (0, 1):     # bottom / top quad -> two triangles
            t = index[key([P[0, 0, dk], P[1, 0, dk], P[1, 1, dk], P[0, 1, dk]])]
            _, ca, cb = faces[t]
            other = cb if ca == a else ca
            dead.add(t)
            faces.append([[P[0, 0, dk], P[1, 0, dk], P[1, 1, dk]], a, other])
            faces.append([[P[0, 0, dk], P[1, 1, dk], P[0, 1, dk]], b, other])
        for quad in ([P[0, 0, 0], P[0, 1, 0], P[0, 1, 1], P[0, 0, 1]], [P[0, 1, 0], P[1, 1, 0], P[1, 1, 1], P[0, 1, 1]]):   # x- and y+ sides go to B
            t = index[key(quad)]
            faces[t][1 if faces[t][1] == a else 2] = b
        faces.append([[P[0, 0, 0], P[1, 1, 0], P[1, 1, 1], P[0, 0, 1]], a, b])   # the cut
    faces = [f for t, f in enumerate(faces) if t not in dead]
    # cell centroids (vertex mean) for the orientation
    acc = np.zeros((n_cells, 3)); cnt = np.zeros(n_cells)
    for loop, ca, cb in faces:
        for c in (ca, cb):
            if c >= 0:
                acc[c] += pts[loop].sum(axis=0); cnt[c] += len(loop)
    cen = acc / cnt[:, None]
    out = []
    for loop, ca, cb in faces:
        if cb >= 0 and cb < ca:
            ca, cb = cb, ca
        p = pts[loop]
        nrm = np.zeros(3)
        for q in range(len(loop)):
            nrm += np.cross(p[q], p[(q + 1) % len(loop)])
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

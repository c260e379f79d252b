"""Host-side numerical check of the formulas behind the opt-in kernel variant k_heavy_hex<., true> (SDFIBM_SYNTH_FACES):
the apex / pyramid volume of a cut hexahedron (reference src/geometrictools.cpp:25-116) evaluated once with the mesh's own face
centres / area vectors and once with the ones the variant forms from the cell's vertices,
    Sf = (p2 - p0) x (p3 - p1) / 2,     apex - Cf = mean(apex - p_i),
in plain Python on sphere-cut cells.  The first evaluation must reproduce the oracle (so the emulation is the algorithm); the second
shows what the variant changes: nothing on integer coordinates (C4), ~ulp(coordinate)/h elsewhere."""
import math

import numpy as np

from oracle.oracle_py import Oracle
from sdfibm_b200.mesh import Mesh
from sdfibm_b200.shapes import make_shape, make_solids

SMALL, TOL = 1e-6, 1e-8


def _line_fraction(a, b):
    if a > 0 and b > 0:
        return 0.0
    if a <= 0 and b <= 0:
        return 1.0
    return -b / (a - b) if a > 0 else -a / (b - a)


def _apex(pts, phi):
    A, pa = pts[0], phi[0]
    B, pb = np.zeros(3), 0.0
    for i in range(1, len(pts)):
        B, pb = pts[i], phi[i]
        if pa * pb <= 0:
            break
    return A - abs(pa) / (SMALL + abs(pa) + abs(pb)) * (A - B)


def _cell_volume(mesh, c, centre, radius, synth):
    vid = mesh.cp[mesh.cp_off[c]:mesh.cp_off[c + 1]]
    P = mesh.points[vid]
    d = np.sqrt(((P - centre) ** 2).sum(axis=1)) - radius
    phi = np.where(np.abs(d) < TOL, -TOL, d)                         # sdf::filter (sdf.h:147-150)
    phi_of = dict(zip(vid.tolist(), phi.tolist()))
    apex = _apex(P, phi)
    vol = 0.0
    for f in mesh.cf[mesh.cf_off[c]:mesh.cf_off[c + 1]]:
        fv = mesh.fp[mesh.fp_off[f]:mesh.fp_off[f + 1]]
        fp = mesh.points[fv]
        ph = [phi_of[int(v)] for v in fv]
        if synth:
            Sf = 0.5 * np.cross(fp[2] - fp[0], fp[3] - fp[1])
            dCf = 0.25 * (((apex - fp[0]) + (apex - fp[1])) + ((apex - fp[2]) + (apex - fp[3])))
            magSf = math.sqrt(float(Sf @ Sf))
        else:
            Sf, dCf = mesh.Sf[f], apex - mesh.Cf[f]
            magSf = math.sqrt(float(Sf @ Sf))
        npos = sum(p > 0 for p in ph)
        if npos == 4:
            eps = 0.0
        elif npos == 0:
            eps = 1.0
        else:
            fap = _apex(fp, ph)
            area = 0.0
            for e in range(4):
                O, A2 = fp[e], fp[(e + 1) % 4]
                cr = np.cross(A2 - O, fap - O)
                area += abs(0.5 * math.sqrt(float(cr @ cr))) * _line_fraction(ph[e], ph[(e + 1) % 4])
            eps = area / magSf
        vol += (1.0 / 3.0) * eps * abs(float(dCf @ Sf))
    return vol


def _run(n, x0, dx, centre, radius):
    mesh = Mesh.hex_block(n, x0=x0, dx=dx)
    shapes = np.array([make_shape("Sphere", radius=radius)])
    S = make_solids(1)
    S[0]["pos"] = centre
    r = Oracle(mesh, False).interact(shapes, S, np.zeros((mesh.n_cells, 3)), 1.0, 1.0)
    off, cells = r["list_off"], r["list_cells"]
    cut = cells[off[1]:off[3]]                                        # CENTER_INSIDE + CENTER_OUTSIDE
    assert len(cut) > 100
    e_mesh = e_synth = 0.0
    for c in cut:
        a_ref = r["Ts"][c]                                            # the unclamped fraction of the one solid
        a_mesh = _cell_volume(mesh, c, np.asarray(centre), radius, False) / mesh.V[c]
        a_syn = _cell_volume(mesh, c, np.asarray(centre), radius, True) / mesh.V[c]
        e_mesh = max(e_mesh, abs(a_mesh - a_ref))
        e_synth = max(e_synth, abs(a_syn - a_ref))
    return e_mesh, e_synth


def test_integer_coordinates_like_c4_change_nothing():
    e_mesh, e_synth = _run((14, 14, 14), (120.0, 121.0, 122.0), (1.0, 1.0, 1.0), (127.3, 127.9, 128.6), 5.0)
    assert e_mesh <= 2e-16                                            # the emulation is the oracle's algorithm
    assert e_synth <= 2e-16                                           # integer vertices, half-integer centres: exact


def test_fine_mesh_near_the_origin_and_far_from_it():
    e_mesh, e_synth = _run((16, 14, 12), (0.37, -0.5, 0.2), (0.1, 0.08, 0.13), (1.13, 0.11, 0.93), 0.47)
    assert e_mesh <= 2e-16 and e_synth <= 1e-13
    e_mesh, e_synth = _run((12, 12, 12), (100.3, 57.7, 33.1), (0.3, 0.3, 0.3), (102.2, 59.4, 35.0), 1.4)
    assert e_mesh <= 2e-16 and 1e-15 < e_synth <= 1e-11              # ulp(100) / 0.3: why the variant must stay restricted

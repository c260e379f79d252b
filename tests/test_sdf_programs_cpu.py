"""Composed shapes: the post-fix SDF programs over the reference's sdf:: namespace (src/libshape/sdf/sdf.h:13-155).
CPU side: the oracle's restatement of the op set, the host façade's interpreter (device_math.cuh compiled for the host — the
code the kernels run) and the hard-coded tag-7 / tag-8 records must agree bit for bit on Circle_Tail / Circle_TwoTail, and the
remaining ops (rot45/60/90/rotth, flipx, I, D, box, ellipse, ellipsoid, half space) must match their sdf.h definitions."""
import os

import numpy as np

import host_cases as hc
from oracle import oracle_py as O
from sdfibm_b200 import hostapi
from sdfibm_b200.shapes import SdfProgram, make_program_shapes, make_shape, make_solids, quat_from_euler_xyz_deg


def _solid(pos=(0.1, -0.05, 0.0), ez=37.0):
    S = make_solids(1)
    S[0]["pos"] = pos
    S[0]["quat"] = quat_from_euler_xyz_deg((0, 0, ez))
    return S


def test_tail_programs_equal_the_hard_coded_records_bit_for_bit(tmp_path):
    r, ratio, th = 0.3, 1.0, 0.1
    ra = (ratio + 1) * 0.5 * r
    tail = SdfProgram().point2d().circle(r).point2d().offset((ra, 0, 0)).rectangle(ra, th).union()
    two = (SdfProgram().point2d().circle(0.2).point2d().rot30().offset((0.25, 0, 0)).rectangle(0.25, 0.05).union()
           .point2d().flipy().rot30().offset((0.25, 0, 0)).rectangle(0.25, 0.05).union())
    recs, ops = make_program_shapes([dict(program=tail, r_out=float(np.hypot(2 * ra, th)) * 1.001, r_in=r, two_d=True),
                                     dict(program=two, r_out=0.6, r_in=0.2, two_d=True)])
    O.set_programs(ops)
    hard = np.array([make_shape("Circle_Tail", radius=r, ratio=ratio, thickness=th), make_shape("Circle_TwoTail", radius=0.2, ratio=1.5, thickness=0.1)])
    pts = np.random.RandomState(0).uniform(-1, 1, size=(50000, 3))
    S = _solid()
    # the façade's Circle_Tail / Circle_TwoTail ARE composed shapes now: their host evaluation runs the interpreter
    meta = dict(on_fluid=0, on_twod=1, gravity=(0, 0, 0))
    path = hc.write_case(tmp_path, meta, [dict(shp_name="tail", mot_name="free", mat_name="heavy", pos=(0, 0, 0))])
    for k, name in enumerate(["tail", "twotail"]):
        a_in, a_phi = O.eval_points(recs[k:k + 1], S[0], pts)
        b_in, b_phi = O.eval_points(hard[k:k + 1], S[0], pts)
        assert np.array_equal(a_in, b_in) and np.array_equal(a_phi, b_phi) and 1000 < a_in.sum() < 20000
        h_in, h_phi = hostapi.shape_eval(path, name, S[0]["pos"], S[0]["quat"], pts)
        want = hc.shape_record(name)
        w_in, w_phi = O.eval_points(np.array([want]), S[0], pts)
        assert np.array_equal(h_in, w_in) and np.array_equal(h_phi, w_phi)
    O.set_programs(None)


def test_every_op_matches_its_sdf_h_definition():
    rng = np.random.RandomState(1)
    pts = rng.uniform(-1.5, 1.5, size=(20000, 3))
    S = make_solids(1)                         # identity pose: the body-frame point is the world point
    x, y, z = pts.T

    def run(prog, two_d=False, com=(0, 0, 0)):
        recs, ops = make_program_shapes([dict(program=prog, r_out=3.0, two_d=two_d, com=com)])
        O.set_programs(ops)
        return O.eval_points(recs, S[0], pts)

    filt = lambda d: np.where(np.abs(d) < 1e-8, -1e-8, d)
    rect = lambda X, Y, ra, rb: (np.hypot(np.maximum(np.abs(X) - ra, 0), np.maximum(np.abs(Y) - rb, 0))
                                  + np.minimum(0.0, np.maximum(np.abs(X) - ra, np.abs(Y) - rb)))
    # rot45 / rot60 / rot90 / rotth / flipx feeding a rectangle (sdf.h:95-122 with the reference's literals)
    for name, (X, Y) in {"rot45": (0.707106781 * (x + y), 0.707106781 * (-x + y)), "rot60": (0.866025404 * y + 0.5 * x, -0.866025404 * x + 0.5 * y),
                         "rot90": (y, -x), "flipx": (-x, y)}.items():
        ins, phi = run(getattr(SdfProgram().point2d(), name)().offset((0.2, 0.1, 0)).rectangle(0.5, 0.3), True)
        Xo, Yo = X - 0.2, Y - 0.1
        assert np.array_equal(ins, (np.abs(Xo) < 0.5) & (np.abs(Yo) < 0.3)), name
        assert np.abs(phi - filt(rect(Xo, Yo, 0.5, 0.3))).max() <= 1e-15, name
    th = 0.4
    ins, phi = run(SdfProgram().point2d().rotth(th).rectangle(0.5, 0.3), True)
    X, Y = x * np.cos(th) + y * np.sin(th), -x * np.sin(th) + y * np.cos(th)
    assert np.array_equal(ins, (np.abs(X) < 0.5) & (np.abs(Y) < 0.3)) and np.abs(phi - filt(rect(X, Y, 0.5, 0.3))).max() <= 1e-15
    # I / D on spheres, with com (sdf::I = max, sdf::D(d1, d2) = max(d1, -d2))
    com = (0.1, 0.0, -0.05)
    P = pts + np.array(com)
    d1 = np.linalg.norm(P, axis=1) - 1.0
    d2 = np.linalg.norm(P - np.array([0.6, 0, 0]), axis=1) - 0.7
    ins, phi = run(SdfProgram().point().sphere(1.0).point().offset((0.6, 0, 0)).sphere(0.7).intersect(), com=com)
    assert np.abs(phi - filt(np.maximum(d1, d2))).max() <= 1e-15 and np.array_equal(ins, ((P ** 2).sum(1) < 1.0) & (((P - [0.6, 0, 0]) ** 2).sum(1) < 0.7 * 0.7))
    ins, phi = run(SdfProgram().point().sphere(1.0).point().offset((0.6, 0, 0)).sphere(0.7).diff(), com=com)
    assert np.abs(phi - filt(np.maximum(d1, -d2))).max() <= 1e-15 and np.array_equal(ins, ((P ** 2).sum(1) < 1.0) & ~(((P - [0.6, 0, 0]) ** 2).sum(1) < 0.7 * 0.7))
    # box / ellipsoid / ellipse / half space against the primitive tags of the same parameters
    for prog, rec in [(SdfProgram().point().box(0.5, 0.3, 0.2), make_shape("Box", radiusa=0.5, radiusb=0.3, radiusc=0.2)),
                      (SdfProgram().point().ellipsoid(0.5, 0.3, 0.2), make_shape("Ellipsoid", radiusa=0.5, radiusb=0.3, radiusc=0.2)),
                      (SdfProgram().point2d().ellipse(0.5, 0.3), make_shape("Ellipse", radiusa=0.5, radiusb=0.3)),
                      (SdfProgram().point2d().circle(0.4), make_shape("Circle", radius=0.4))]:
        a = run(prog)
        b = O.eval_points(np.array([rec]), S[0], pts)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    ins, phi = run(SdfProgram().point().halfspace().point().sphere(1.0).intersect())     # a half ball
    assert np.array_equal(ins, (y < 0) & ((pts ** 2).sum(1) < 1.0))
    O.set_programs(None)

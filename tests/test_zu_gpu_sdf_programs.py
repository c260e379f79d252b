"""Composed shapes on the device (row g1): SDFIBM_SHAPE_PROGRAM records evaluated by the interact kernels — parity with the oracle's
restatement of sdf.h (lists bit-exact, fields 1e-12), the two tail shapes as programs against their hard-coded tags, and the
validation of malformed programs."""
import numpy as np
import pytest

from oracle import oracle_py as O
from oracle.oracle_py import Oracle
from sdfibm_b200 import cases
from sdfibm_b200.capi import SdfibmError
from sdfibm_b200.context import Context
from sdfibm_b200.mesh import Mesh
from sdfibm_b200.shapes import SdfProgram, make_program_shapes, make_shape, make_solids, quat_from_euler_xyz_deg
from test_gpu_parity import check_parity

pytestmark = pytest.mark.gpu


def _run(case, ops, cell_slots=8):
    O.set_programs(ops)
    try:
        o = Oracle(case["mesh"], case["two_d"])
        ref = o.interact(case["shapes"], case["solids"], case["U"], case["dt"], case["rhof"])
        ctx = Context(0, cell_slots=cell_slots)
        ctx.set_mesh(case["mesh"], case["two_d"])
        ctx.set_shape_programs(ops)
        ctx.set_shapes(case["shapes"])
        got = ctx.interact(case["solids"], case["U"], case["dt"], case["rhof"])
        check_parity(case, o, ref, ctx, got)
        pairs = sum(ctx.candidate_counts())
        ctx.close()
        return got, pairs
    finally:
        O.set_programs(None)


def test_composed_2d_shapes_match_the_oracle():
    """A gear (circle + three rotated bars - hub), a capsule and a plain program circle, mixed with primitive tags in one table."""
    gear = (SdfProgram().point2d().circle(0.25)
            .point2d().rectangle(0.45, 0.06).union()
            .point2d().rot60().rectangle(0.45, 0.06).union()
            .point2d().rotth(2.0943951023931953).rectangle(0.45, 0.06).union()
            .point2d().circle(0.1).diff())
    capsule = (SdfProgram().point2d().offset((-0.2, 0, 0)).circle(0.12).point2d().offset((0.2, 0, 0)).circle(0.12).union()
               .point2d().rectangle(0.2, 0.12).union())
    prog, ops = make_program_shapes([dict(program=gear, r_out=0.46, r_in=0.0, two_d=True), dict(program=capsule, r_out=0.33, r_in=0.12, two_d=True, com=(0.02, -0.01, 0.0)),
                                     dict(program=SdfProgram().point2d().circle(0.2), r_out=0.2, r_in=0.2, two_d=True)])
    shapes = np.concatenate([prog, np.array([make_shape("Circle", radius=0.2), make_shape("Plane")])])
    mesh = Mesh.hex_block((80, 80, 1), x0=(-2.0, -2.0, -0.5), dx=(0.05, 0.05, 1.0))
    rng = np.random.RandomState(4)
    S = make_solids(9)
    S["pos"][:8, :2] = rng.uniform(-1.5, 1.5, size=(8, 2))
    S["shape"][:8] = [0, 1, 2, 3, 0, 1, 0, 1]
    for i in range(8):
        S[i]["quat"] = quat_from_euler_xyz_deg((0, 0, float(rng.uniform(-180, 180))))
    S[8]["pos"] = (0.0, -1.8, 0.0); S[8]["shape"] = 4
    S["vel"][:8] = 0.1 * rng.standard_normal((8, 3)); S["vel"][:, 2] = 0
    S["omega"][:8, 2] = rng.standard_normal(8)
    case = dict(name="programs2d", mesh=mesh, two_d=True, shapes=shapes, solids=S, U=cases.taylor_green(mesh.cc, 4.0), dt=1e-3, rhof=1.1)
    got, pairs = _run(case, ops)
    assert pairs > 1000
    # solids 2 (program circle) and 3 (tag circle) have the same radius: equal cell counts as lists are bit-exact per solid


def test_composed_3d_shape_on_a_box_mesh_and_a_skewed_mesh():
    lens = SdfProgram().point().offset((-0.8, 0, 0)).sphere(1.6).point().offset((0.8, 0, 0)).sphere(1.6).intersect()      # two-sphere lens
    cut = SdfProgram().point().box(1.2, 0.9, 0.7).point().sphere(0.6).diff()                                             # box minus ball
    prog, ops = make_program_shapes([dict(program=lens, r_out=1.4, r_in=0.75), dict(program=cut, r_out=1.7, r_in=0.0)])
    shapes = np.concatenate([prog, np.array([make_shape("Sphere", radius=1.1)])])
    rng = np.random.RandomState(9)
    for mesh in (Mesh.hex_block((20, 20, 20), (0, 0, 0), (0.5, 0.5, 0.5)), cases.case_mixed3d(n=12)["mesh"]):
        L = float(mesh.bounds_max[0] - mesh.bounds_min[0])
        S = make_solids(7)
        S["pos"] = rng.uniform(0.15 * L, 0.85 * L, size=(7, 3))
        S["shape"] = [0, 1, 2, 0, 1, 0, 1]
        for i in range(7):
            S[i]["quat"] = quat_from_euler_xyz_deg(tuple(rng.uniform(-90, 90, size=3)))
        S["vel"] = 0.1 * rng.standard_normal((7, 3))
        S["omega"] = 0.2 * rng.standard_normal((7, 3))
        case = dict(name="programs3d", mesh=mesh, two_d=False, shapes=shapes, solids=S, U=cases.taylor_green(mesh.cc, L), dt=1e-3, rhof=1.0)
        got, pairs = _run(case, ops)
        assert pairs > 100


def test_tail_programs_give_the_same_fields_as_the_hard_coded_tags(m1_points):
    case = cases.case_g1(m1_points)          # G1: Circle_Tail x3 among its 14 solids (tag 7)
    r, ratio, th = 0.3, 1.0, 0.1
    ra = (ratio + 1) * 0.5 * r
    tail = SdfProgram().point2d().circle(r).point2d().offset((ra, 0, 0)).rectangle(ra, th).union()
    # certified outer radius: the far corners of the tail box [0, 2 ra] x [-rb, rb]
    prog, ops = make_program_shapes([dict(program=tail, r_out=float(np.hypot(2 * ra, th)) * 1.001, r_in=r, two_d=True)])
    shapes = case["shapes"].copy()
    assert shapes[1]["tag"] == 7
    shapes[1] = prog[0]
    ctx = Context(0, cell_slots=8)
    ctx.set_mesh(case["mesh"], True)
    ctx.set_shapes(case["shapes"])
    a = ctx.interact(case["solids"], case["U"], case["dt"], case["rhof"])
    la = ctx.candidate_lists()
    ctx.set_shape_programs(ops)
    ctx.set_shapes(shapes)
    b = ctx.interact(case["solids"], case["U"], case["dt"], case["rhof"])
    lb = ctx.candidate_lists()
    assert np.array_equal(la[0], lb[0]) and np.array_equal(la[1], lb[1])
    for k in ("As", "Fs", "Ts", "Ct"):
        assert np.array_equal(a[k], b[k]), k
    ctx.close()


def test_malformed_programs_are_refused():
    ctx = Context(0)
    good = SdfProgram().point().sphere(1.0)
    recs, ops = make_program_shapes([dict(program=good, r_out=1.0)])
    with pytest.raises(SdfibmError, match="op table"):
        ctx.set_shapes(recs)                                   # programs not set yet
    for bad, msg in [(SdfProgram().sphere(1.0), "without a point"), (SdfProgram().point().sphere(1.0).union(), "without two values"),
                     (SdfProgram().point().point().sphere(1.0), "exactly one value"), (SdfProgram().point().sphere(1.0).point().sphere(2.0), "exactly one value")]:
        r2, o2 = make_program_shapes([dict(program=bad, r_out=1.0)])
        ctx.set_shape_programs(o2)
        with pytest.raises(SdfibmError, match=msg):
            ctx.set_shapes(r2)
    r3, o3 = make_program_shapes([dict(program=good, r_out=1.0)])
    r3[0]["p"][2] = 0.0
    ctx.set_shape_programs(o3)
    with pytest.raises(SdfibmError, match="certified radii"):
        ctx.set_shapes(r3)
    ctx.close()

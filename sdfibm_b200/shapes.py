"""Shape / solid record builders: the members each reference IShape subclass derives in its constructor
(reference src/libshape/*.h) lowered to the POD records of include/sdfibm_b200.h."""
from __future__ import annotations

import math

import numpy as np

from .capi import SDF_OP_DTYPE, SDF_OPS, SHAPE_DTYPE, SHAPE_PROGRAM_TAG, SHAPE_TAGS, SOLID_DTYPE


def make_shape(type_name: str, com=(0.0, 0.0, 0.0), **k) -> np.ndarray:
    """One sdfibm_shape_t from the solidDict keys of the shape (`radius`, `radiusa`, `ratio`, ...)."""
    if type_name not in SHAPE_TAGS:
        raise ValueError(f"Cannot create unrecognized object: {type_name}")  # genericfactory.h:39
    s = np.zeros((), dtype=SHAPE_DTYPE)
    s["tag"] = SHAPE_TAGS[type_name]
    s["finite"] = 1
    s["com"] = com
    p = np.zeros(8)
    if type_name in ("Circle", "Sphere"):
        r = float(k["radius"])
        p[0], p[1] = r, r * r
        s["radiusB"] = r
    elif type_name == "Ellipse":
        a, b = float(k["radiusa"]), float(k["radiusb"])
        p[0], p[1], p[2], p[3] = a, b, 1.0 / (a * a), 1.0 / (b * b)
        s["radiusB"] = max(a, b)
    elif type_name == "Ellipsoid":
        a, b, c = float(k["radiusa"]), float(k["radiusb"]), float(k["radiusc"])
        p[0:6] = a, b, c, 1.0 / (a * a), 1.0 / (b * b), 1.0 / (c * c)
        s["radiusB"] = max(max(a, b), c)
        s["com"] = (0.0, 0.0, 0.0)  # ellipsoid.h never reads `com`
    elif type_name == "Rectangle":
        a, b = float(k["radiusa"]), float(k["radiusb"])
        p[0], p[1] = a, b
        s["radiusB"] = max(a, b)
    elif type_name == "Box":
        a, b, c = float(k["radiusa"]), float(k["radiusb"]), float(k["radiusc"])
        p[0:3] = a, b, c
        s["radiusB"] = max(max(a, b), c)
    elif type_name in ("Circle_Tail", "Circle_TwoTail"):
        r, ratio, th = float(k["radius"]), float(k["ratio"]), float(k["thickness"])
        ra = (ratio + 1) * 0.5 * r
        p[0], p[1], p[2] = r, r * r, ra
        p[3] = th if type_name == "Circle_Tail" else th * 0.5
        s["radiusB"] = 2 * ra
    elif type_name == "Plane":
        s["finite"] = 0
        s["radiusB"] = 0.0
    s["p"] = p
    return s


def quat_from_euler_xyz_deg(euler_deg):
    """Foam::quaternion(XYZ, angles): q_x(a_x) * q_y(a_y) * q_z(a_z), angles in degrees as in solidDict
    (reference src/solidcloud.cpp:177, src/solid.h:76-80)."""
    ax, ay, az = [float(e) * math.pi / 180.0 for e in euler_deg]

    def qaxis(axis, th):
        return (math.cos(0.5 * th),) + tuple(math.sin(0.5 * th) * a for a in axis)

    def qmul(a, b):
        w1, x1, y1, z1 = a
        w2, x2, y2, z2 = b
        return (
            w1 * w2 - (x1 * x2 + y1 * y2 + z1 * z2),
            w1 * x2 + w2 * x1 + (y1 * z2 - z1 * y2),
            w1 * y2 + w2 * y1 + (z1 * x2 - x1 * z2),
            w1 * z2 + w2 * z1 + (x1 * y2 - y1 * x2),
        )

    q = qaxis((1, 0, 0), ax)
    q = qmul(q, qaxis((0, 1, 0), ay))
    q = qmul(q, qaxis((0, 0, 1), az))
    return q


def make_solids(n: int) -> np.ndarray:
    s = np.zeros(n, dtype=SOLID_DTYPE)
    s["quat"][:, 0] = 1.0
    return s


class SdfProgram:
    """Builder of a composed shape's post-fix program over the reference's sdf:: namespace (src/libshape/sdf/sdf.h; the op set is
    documented at sdfibm_sdf_op_t in include/sdfibm_b200.h).  Each method appends one op and returns self:

        tail = (SdfProgram().point2d().circle(r)                                   # d1 = sdf::circle(p2d, r)
                .point2d().offset((ra, 0, 0)).rectangle(ra, rb)                    # d2 = sdf::rectangle(sdf::offset(p2d, ..), ra, rb)
                .union())                                                          # sdf::U({d1, d2})
    """

    def __init__(self):
        self.ops = []

    def _add(self, name, *a):
        o = np.zeros((), dtype=SDF_OP_DTYPE)
        o["op"] = SDF_OPS[name]
        o["a"][: len(a)] = a
        self.ops.append(o)
        return self

    def point(self): return self._add("POINT")
    def point2d(self): return self._add("POINT_2D")
    def offset(self, v): return self._add("OFFSET", *[float(x) for x in v])
    def rot30(self): return self._add("ROT30")
    def rot45(self): return self._add("ROT45")
    def rot60(self): return self._add("ROT60")
    def rot90(self): return self._add("ROT90")
    def rotth(self, th): return self._add("ROTTH", float(th))
    def flipx(self): return self._add("FLIPX")
    def flipy(self): return self._add("FLIPY")
    def circle(self, r): return self._add("CIRCLE", float(r), float(r) * float(r))
    sphere = circle
    def rectangle(self, ra, rb): return self._add("RECTANGLE", float(ra), float(rb))
    def box(self, ra, rb, rc): return self._add("BOX", float(ra), float(rb), float(rc))
    def ellipse(self, a, b): return self._add("ELLIPSE", 1.0 / (a * a), 1.0 / (b * b))
    def ellipsoid(self, a, b, c): return self._add("ELLIPSOID", 1.0 / (a * a), 1.0 / (b * b), 1.0 / (c * c))
    def halfspace(self): return self._add("HALFSPACE")
    def union(self): return self._add("UNION")
    def intersect(self): return self._add("INTERSECT")
    def diff(self): return self._add("DIFF")


def make_program_shapes(programs):
    """(shape records, op table) of composed shapes: `programs` = list of dict(program=SdfProgram, r_out=, r_in=0.0, two_d=False,
    com=(0, 0, 0), radiusB=None).  r_out / r_in are the certified radii about the body origin the binning needs."""
    ops, recs = [], []
    for d in programs:
        s = np.zeros((), dtype=SHAPE_DTYPE)
        s["tag"] = SHAPE_PROGRAM_TAG
        s["finite"] = 1
        s["com"] = d.get("com", (0.0, 0.0, 0.0))
        s["radiusB"] = d.get("radiusB") or d["r_out"]
        p = np.zeros(8)
        p[0], p[1], p[2], p[3], p[4] = len(ops), len(d["program"].ops), d["r_out"], d.get("r_in", 0.0), 1.0 if d.get("two_d") else 0.0
        s["p"] = p
        recs.append(s)
        ops.extend(d["program"].ops)
    return np.array(recs, dtype=SHAPE_DTYPE), np.array(ops, dtype=SDF_OP_DTYPE)

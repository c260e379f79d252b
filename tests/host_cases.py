"""Shared set-ups of the host-façade tests: a solidDict written from Python data plus the same data for the oracle."""
import math
import os

import numpy as np

from oracle import host_oracle as ho
from sdfibm_b200 import hostapi
from sdfibm_b200.shapes import make_shape, quat_from_euler_xyz_deg

SHAPES = {
    "circ": dict(type="Circle", radius=0.15),
    "sph": dict(type="Sphere", radius=0.3, com=(0.01, -0.02, 0.0)),
    "ell": dict(type="Ellipse", radiusa=0.3, radiusb=0.15),
    "elo": dict(type="Ellipsoid", radiusa=0.5, radiusb=0.45, radiusc=0.4),
    "rect": dict(type="Rectangle", radiusa=0.3, radiusb=0.2),
    "box": dict(type="Box", radiusa=0.26, radiusb=0.17, radiusc=0.31),
    "tail": dict(type="Circle_Tail", radius=0.3, ratio=1.0, thickness=0.1),
    "twotail": dict(type="Circle_TwoTail", radius=0.2, ratio=1.5, thickness=0.1),
    "plane": dict(type="Plane"),
    "circle_tc": dict(type="Circle", radius=0.3),      # examples/taylor_couette/solidDict: circle1
}
MOTIONS = {
    "mask": dict(type="Motion01Mask", mask="b110001"),
    "spin": dict(type="Motion000002", period=2.0),
    "spinfree": dict(type="Motion110002", period=3.0),
    "const": dict(type="Motion222000", u=0.1, v=-0.2, w=0.05),
    "sine": dict(type="MotionSineDirectional", amplitude=0.3, period=1.5, direction=(0.6, 0.8, 0.0)),
    "rotor": dict(type="MotionRotor", period=4.0, radius=0.7, theta0=0.3, selfom=1.1),
    "gate": dict(type="MotionOpenClose"),
}
FORCES = {
    "push": dict(type="Constant", force=(0.1, 0.2, -0.3), torque=(0.01, 0.0, 0.02)),
    "spring": dict(type="Spring", pivot=(0.0, 1.0, 0.0), k=5.0, l=0.4),
    "mag": dict(type="Magnetic", direction=(1.0, 0.0, 0.0), A=0.7, w=2.0),
}
MATERIALS = {"heavy": dict(type="General", rho=3.0), "light": dict(type="General", rho=1.2)}


def shape_record(name):
    d = dict(SHAPES[name])
    t = d.pop("type")
    return make_shape(t, **d)


def mass_props(name):
    """(volume, volumeINV, moi_inv diagonal) as the reference's shape constructors compute them."""
    d = SHAPES[name]
    t = d["type"]
    if t in ("Circle", "Circle_Tail", "Circle_TwoTail"):
        r2 = d["radius"] * d["radius"]
        vol = math.pi * r2
        moi = [0.5 * vol * r2] * 3
    elif t == "Sphere":
        r2 = d["radius"] * d["radius"]
        vol = 4.0 / 3.0 * math.pi * r2 * d["radius"]
        moi = [0.4 * vol * r2] * 3
    elif t == "Ellipse":
        a, b = d["radiusa"], d["radiusb"]
        vol = math.pi * a * b
        moi = [0.25 * vol * (a * a + b * b)] * 3
    elif t == "Ellipsoid":
        a, b, c = d["radiusa"], d["radiusb"], d["radiusc"]
        vol = 4.0 / 3.0 * math.pi * a * b * c
        moi = [0.2 * vol * (b * b + c * c), 0.2 * vol * (a * a + c * c), 0.2 * vol * (a * a + b * b)]
    elif t == "Rectangle":
        a, b = d["radiusa"], d["radiusb"]
        vol = 4.0 * a * b
        moi = [1.0 / 3.0 * vol * (a * a + b * b)] * 3
    elif t == "Box":
        a, b, c = d["radiusa"], d["radiusb"], d["radiusc"]
        vol = 8.0 * a * b * c
        moi = [1.0 / 3.0 * vol * (b * b + c * c), 1.0 / 3.0 * vol * (a * a + c * c), 1.0 / 3.0 * vol * (b * b + a * a)]
    elif t == "Plane":
        return 0.0, 0.0, [1.0, 1.0, 1.0]
    return vol, 1.0 / vol, [1.0 / m for m in moi]


def write_case(tmpdir, meta, solids, shapes=None, motions=None, forces=None, materials=None):
    shapes = shapes or SHAPES
    return hostapi.write_solid_dict(os.path.join(str(tmpdir), "solidDict"), meta, shapes, motions or MOTIONS, materials or MATERIALS,
                                    solids, forces if forces is not None else FORCES)


def oracle_solids(solids, motions=None, forces=None, materials=None):
    motions, forces, materials = motions or MOTIONS, forces if forces is not None else FORCES, materials or MATERIALS
    out = []
    for s in solids:
        vol, vinv, minv = mass_props(s["shp_name"])
        q = quat_from_euler_xyz_deg(s.get("euler", (0, 0, 0)))
        out.append(ho.SolidState(s["pos"], q, s.get("vel", (0, 0, 0)), s.get("omega", (0, 0, 0)), vol, vinv, minv,
                                 materials[s["mat_name"]]["rho"],
                                 None if s["mot_name"] == "free" else motions[s["mot_name"]],
                                 forces[s["for_name"]] if "for_name" in s else None))
    return out


def shape_table(solids, shapes=None):
    """Shape table in solidDict order + per-solid index, like SolidCloud::buildShapeTable."""
    shapes = shapes or SHAPES
    names = list(shapes)
    table = np.array([shape_record(n) if shapes is SHAPES else None for n in names])
    return table, [names.index(s["shp_name"]) for s in solids]


def state_arrays(osolids):
    return (np.array([s.x for s in osolids]), np.array([(s.q[0],) + s.q[1] for s in osolids]), np.array([s.v for s in osolids]),
            np.array([s.om for s in osolids]))

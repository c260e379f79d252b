"""CPU restatement of the reference's rigid-body side (TEST INFRASTRUCTURE ONLY — see oracle.cpp's header).

Plain Python floats, operation order of the reference:
  Solid::move / applyForcer / addMidFluidForceAndTorque / storeOldForce      reference src/solid.h:131-202
  SolidCloud::evolve / addMidEnvironment / solidSolidInteract                 reference src/solidcloud.cpp:466-562
  motions                                                                     reference src/libmotion/*.h
  forcers                                                                     reference src/libforcer/{constant,spring,magnetic}.h
  quaternion arithmetic (OpenFOAM quaternionI.H, not vendored)                restated as in oracle.cpp
Parity unpinned by any reference artefact (the reference ships no trajectory fixtures): these functions pin the host
façade (sdfibm_b200/host) against an independent statement of the same formulas.
"""
from __future__ import annotations

import math

import numpy as np


# ---- vector / quaternion helpers (tuples of floats) ---------------------------------------------
def vadd(a, b): return (a[0] + b[0], a[1] + b[1], a[2] + b[2])
def vsub(a, b): return (a[0] - b[0], a[1] - b[1], a[2] - b[2])
def vscale(s, a): return (s * a[0], s * a[1], s * a[2])
def vmuls(a, s): return (a[0] * s, a[1] * s, a[2] * s)
def dot(a, b): return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
def cross(a, b): return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])
def mag(a): return math.sqrt(dot(a, a))


def qmul(a, b):
    (w1, v1), (w2, v2) = a, b
    return (w1 * w2 - dot(v1, v2), vadd(vadd(vscale(w1, v2), vscale(w2, v1)), cross(v1, v2)))


def qtransform(q, u):
    w, v = q
    m = (-dot(v, u), vadd(vscale(w, u), cross(v, u)))
    return qmul(m, (w, (-v[0], -v[1], -v[2])))[1]


def qR(q):
    w, (x, y, z) = q
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    txy, twz, txz, twy, tyz, twx = 2 * x * y, 2 * w * z, 2 * x * z, 2 * w * y, 2 * y * z, 2 * w * x
    return [[w2 + x2 - y2 - z2, txy - twz, txz + twy], [txy + twz, w2 - x2 + y2 - z2, tyz - twx],
            [txz - twy, tyz + twx, w2 - x2 - y2 + z2]]


def mat_mul(a, b):
    return [[a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j] for j in range(3)] for i in range(3)]


def mat_vec(a, v):
    return tuple(a[i][0] * v[0] + a[i][1] * v[1] + a[i][2] * v[2] for i in range(3))


def transpose(a):
    return [[a[j][i] for j in range(3)] for i in range(3)]


def euler_xyz(q):
    w, (x, y, z) = q
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    return (math.atan2(2 * (w * x - y * z), w2 - x2 - y2 + z2), math.asin(2 * (x * z + w * y)),
            math.atan2(2 * (w * z - x * y), w2 + x2 - y2 - z2))


# ---- plugins ---------------------------------------------------------------------------------------
def motion_constraint(m, time, v, om):
    """m = dict(type=..., keys of the solidDict entry); returns (v, omega)."""
    if m is None:
        return v, om
    t = m["type"]
    if t == "Motion01Mask":
        k = m["mask"]
        vm = tuple(0.0 if k[i] == "0" else 1.0 for i in (1, 2, 3))
        omk = tuple(0.0 if k[i] == "0" else 1.0 for i in (4, 5, 6))
        return (v[0] * vm[0], v[1] * vm[1], v[2] * vm[2]), (om[0] * omk[0], om[1] * omk[1], om[2] * omk[2])
    if t == "Motion000002":
        return (0.0, 0.0, 0.0), (0.0, 0.0, 2 * math.pi / m["period"])
    if t == "Motion110002":
        return (v[0], v[1], 0.0), (0.0, 0.0, 2 * math.pi / m["period"])
    if t == "Motion222000":
        return (m["u"], m["v"], m["w"]), (0.0, 0.0, 0.0)
    if t == "MotionSineDirectional":
        w = 2 * math.pi / m["period"]
        return vmuls(m["direction"], 1.0) and vscale(m["amplitude"] * w * math.cos(w * time), m["direction"]), (0.0, 0.0, 0.0)
    if t == "MotionRotor":
        w = 2 * math.pi / m["period"]
        return ((-m["radius"] * w * math.sin(w * time + m["theta0"]), m["radius"] * w * math.cos(w * time + m["theta0"]), 0.0),
                (0.0, 0.0, m["selfom"]))
    if t == "MotionOpenClose":
        tt = math.fmod(time, 5.0)
        vy = -1.0 if 1 < tt < 2 else (1.0 if 3 < tt < 4 else 0.0)
        return (0.0, vy, 0.0), (0.0, 0.0, 0.0)
    raise ValueError(t)


def forcer_generate(f, time, x, v, q, om):
    t = f["type"]
    if t == "Constant":
        return tuple(f["force"]), tuple(f["torque"])
    if t == "Spring":
        r = vsub(x, f["pivot"])
        force = (0.0, 0.0, 0.0)
        if mag(r) > 1e-6:
            force = vmuls(vscale(-f["k"], r), 1.0 - f["l"] / mag(r))
        return force, (0.0, 0.0, 0.0)
    if t == "Magnetic":
        B = vscale(f["A"] * math.cos(f["w"] * time), f["direction"])
        m = qtransform((q[0], (-q[1][0], -q[1][1], -q[1][2])), vscale(1.0, (0.0, 0.0, 1.0)))
        return (0.0, 0.0, 0.0), cross(m, B)
    raise ValueError(t)


class SolidState:
    def __init__(self, pos, quat, vel, omega, volume, volume_inv, moi_inv_diag, rho, motion=None, forcer=None):
        self.x, self.q, self.v, self.om = tuple(pos), (quat[0], tuple(quat[1:])), tuple(vel), tuple(omega)
        self.mass = volume * rho
        self.mass_inv = volume_inv / rho
        self.moi_inv = [[moi_inv_diag[i] / rho if i == j else 0.0 / rho for j in range(3)] for i in range(3)]
        self.rho = rho
        self.motion, self.forcer = motion, forcer
        z = (0.0, 0.0, 0.0)
        self.force = self.torque = z
        self.ff = self.ft = self.ff_old = self.ft_old = z
        self.gf = self.gt = self.gf_old = self.gt_old = z
        self.first_fluid = self.first_forcer = True

    def move(self, time, dt):
        v_old, om_old = self.v, self.om
        self.v = vadd(self.v, vmuls(vmuls(self.force, self.mass_inv), dt))
        R = qR(self.q)
        miw = mat_mul(mat_mul(R, self.moi_inv), transpose(R))
        self.om = vadd(self.om, vmuls(mat_vec(miw, self.torque), dt))
        self.v, self.om = motion_constraint(self.motion, time, self.v, self.om)
        self.x = vadd(self.x, vmuls(vscale(0.5, vadd(self.v, v_old)), dt))
        qw = (0.0, vscale(0.5, vadd(self.om, om_old)))
        inc = qmul((0.5 * qw[0], vscale(0.5, qw[1])), self.q)
        inc = (inc[0] * dt, vmuls(inc[1], dt))
        q = (self.q[0] + inc[0], vadd(self.q[1], inc[1]))
        n = math.sqrt(q[0] * q[0] + dot(q[1], q[1]))
        self.q = (q[0] / n, (q[1][0] / n, q[1][1] / n, q[1][2] / n))


def evolve(solids, time, dt, n_subiter, gravity, rhof, collide=None):
    """SolidCloud::evolve.  collide(solids) -> [6N] array of collision (force, torque) or None."""
    dt_sub = dt / n_subiter
    for _ in range(n_subiter):
        for s in solids:
            s.force = s.torque = (0.0, 0.0, 0.0)
        for s in solids:
            if s.forcer is not None:
                s.gf, s.gt = forcer_generate(s.forcer, time, s.x, s.v, s.q, s.om)
                if s.first_forcer:
                    s.gf_old, s.gt_old, s.first_forcer = s.gf, s.gt, False
                s.force = vadd(s.force, vsub(vscale(1.5, s.gf), vscale(0.5, s.gf_old)))
                s.torque = vadd(s.torque, vsub(vscale(1.5, s.gt), vscale(0.5, s.gt_old)))
        for s in solids:
            if s.first_fluid:
                s.ff_old, s.ft_old, s.first_fluid = s.ff, s.ft, False
            s.force = vadd(s.force, vsub(vscale(1.5, s.ff), vscale(0.5, s.ff_old)))
            s.torque = vadd(s.torque, vsub(vscale(1.5, s.ft), vscale(0.5, s.ft_old)))
        for s in solids:
            gp = vscale((s.rho - rhof) / s.rho, gravity)
            s.force = vadd(s.force, vscale(s.mass, gp))
        if collide is not None:
            ft = collide(solids)
            if ft is not None:
                for i, s in enumerate(solids):
                    s.force = vadd(s.force, tuple(ft[i, :3]))
                    s.torque = vadd(s.torque, tuple(ft[i, 3:]))
        for s in solids:
            s.move(time, dt_sub)
    for s in solids:
        s.ff_old, s.ft_old, s.gf_old, s.gt_old = s.ff, s.ft, s.gf, s.gt


def records(solids, shape_index):
    from sdfibm_b200.capi import SOLID_DTYPE
    out = np.zeros(len(solids), dtype=SOLID_DTYPE)
    for i, s in enumerate(solids):
        out[i]["pos"] = s.x
        out[i]["quat"] = (s.q[0],) + s.q[1]
        out[i]["vel"] = s.v
        out[i]["omega"] = s.om
        out[i]["shape"] = shape_index[i]
    return out

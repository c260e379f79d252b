"""ctypes wrapper of oracle/_ref/libsdfibm_ref.so — TEST INFRASTRUCTURE ONLY.

The library is the reference's own CellEnumerator / GeometricTools / shape classes, compiled unmodified from /root/reference
(oracle/Makefile, target `ref`); it exists only where the reference tree does (this container, not the GPU box).  It is used to
validate the restatement in oracle.cpp; it is never on a product path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libsdfibm_ref.so")


def available() -> bool:
    return os.path.exists(LIB_PATH)


class _RefMesh(C.Structure):
    _fields_ = [("n_cells", C.c_int32), ("n_points", C.c_int32), ("n_faces", C.c_int32)] + \
               [(k, C.c_void_p) for k in ("points", "cc", "V", "Cf", "Sf", "cp_off", "cp", "cf_off", "cf", "fp_off", "fp", "nb_off", "nb")]


def shape_dict_text(type_name: str, com=(0.0, 0.0, 0.0), **k) -> str:
    """The solidDict entries of one shape, as the reference's constructors read them (src/libshape/*.h)."""
    s = f"type {type_name};\n"
    for key, v in k.items():
        s += f"{key} {float(v)!r};\n"
    if any(c != 0.0 for c in com):
        s += f"com ({float(com[0])!r} {float(com[1])!r} {float(com[2])!r});\n"
    return s


def ref_interact(mesh, dict_texts, pos, quat, seeds, two_d: bool):
    """Candidate lists (offsets[3n+1], cells) and the clipped volume-fraction field of the reference's own code."""
    lib = C.CDLL(LIB_PATH)
    lib.ref_interact.restype = C.c_int64
    keep = [np.ascontiguousarray(a) for a in (mesh.points, mesh.cc, mesh.V, mesh.Cf, mesh.Sf, mesh.cp_off, mesh.cp, mesh.cf_off, mesh.cf,
                                              mesh.fp_off, mesh.fp, mesh.nb_off, mesh.nb)]
    m = _RefMesh(mesh.n_cells, mesh.n_points, mesh.n_faces, *[a.ctypes.data for a in keep])
    n = len(dict_texts)
    texts = (C.c_char_p * n)(*[t.encode() for t in dict_texts])
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    quat = np.ascontiguousarray(quat, dtype=np.float64)
    seeds = np.ascontiguousarray(seeds, dtype=np.int32)
    cap = 8 * mesh.n_cells + 64
    off = np.zeros(3 * n + 1, dtype=np.int32)
    cells = np.zeros(cap, dtype=np.int32)
    As = np.zeros(mesh.n_cells)
    rc = lib.ref_interact(C.byref(m), n, texts, C.c_void_p(pos.ctypes.data), C.c_void_p(quat.ctypes.data), C.c_void_p(seeds.ctypes.data),
                          int(bool(two_d)), C.c_void_p(off.ctypes.data), C.c_void_p(cells.ctypes.data), C.c_int64(cap),
                          C.c_void_p(As.ctypes.data))
    if rc < 0:
        raise RuntimeError(f"ref_interact failed ({rc})")
    return off, cells[:rc].copy(), As


def ref_collide(bounds_min, bounds_max, delta, dict_texts, pos, quat):
    """Pairs in UGrid::generateCollisionPairs order and the accumulated contact forces [n, 6] of the reference's libcollision."""
    lib = C.CDLL(LIB_PATH)
    lib.ref_collide.restype = C.c_int64
    n = len(dict_texts)
    texts = (C.c_char_p * n)(*[t.encode() for t in dict_texts])
    bmin = np.ascontiguousarray(bounds_min, dtype=np.float64)
    bmax = np.ascontiguousarray(bounds_max, dtype=np.float64)
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    quat = np.ascontiguousarray(quat, dtype=np.float64)
    cap = 64 * n + 64
    pairs = np.zeros((cap, 2), dtype=np.int32)
    ft = np.zeros((n, 6))
    rc = lib.ref_collide(C.c_void_p(bmin.ctypes.data), C.c_void_p(bmax.ctypes.data), C.c_double(delta), n, texts,
                         C.c_void_p(pos.ctypes.data), C.c_void_p(quat.ctypes.data), C.c_void_p(pairs.ctypes.data), C.c_int64(cap),
                         C.c_void_p(ft.ctypes.data))
    if rc < 0:
        raise RuntimeError(f"ref_collide failed ({rc})")
    return pairs[:rc].copy(), ft


def dict_text_from_record(rec) -> str:
    """solidDict entries that make the reference's constructor rebuild a lowered sdfibm_shape_t record (sdfibm_b200/shapes.py)."""
    names = {0: "Plane", 1: "Circle", 2: "Sphere", 3: "Ellipse", 4: "Ellipsoid", 5: "Rectangle", 6: "Box", 7: "Circle_Tail", 8: "Circle_TwoTail"}
    t = names[int(rec["tag"])]
    p = [float(x) for x in rec["p"]]
    k = {}
    if t in ("Circle", "Sphere"):
        k = dict(radius=p[0])
    elif t in ("Ellipse", "Rectangle"):
        k = dict(radiusa=p[0], radiusb=p[1])
    elif t in ("Ellipsoid", "Box"):
        k = dict(radiusa=p[0], radiusb=p[1], radiusc=p[2])
    elif t in ("Circle_Tail", "Circle_TwoTail"):
        k = dict(radius=p[0], ratio=2.0 * p[2] / p[0] - 1.0, thickness=p[3] if t == "Circle_Tail" else 2.0 * p[3])
    return shape_dict_text(t, com=tuple(float(x) for x in rec["com"]), **k)


class Reference:
    """interact() of the reference's own compiled classes on one mesh (the loop of src/solidcloud.cpp:361-464 around them)."""

    def __init__(self, mesh):
        self._lib = C.CDLL(LIB_PATH)
        self._lib.ref_create.restype = C.c_void_p
        self._lib.ref_interact_full.restype = C.c_int64
        self._keep = [np.ascontiguousarray(a) for a in (mesh.points, mesh.cc, mesh.V, mesh.Cf, mesh.Sf, mesh.cp_off, mesh.cp, mesh.cf_off,
                                                        mesh.cf, mesh.fp_off, mesh.fp, mesh.nb_off, mesh.nb)]
        m = _RefMesh(mesh.n_cells, mesh.n_points, mesh.n_faces, *[a.ctypes.data for a in self._keep])
        self.mesh = mesh
        self._h = C.c_void_p(self._lib.ref_create(C.byref(m)))
        if not self._h:
            raise RuntimeError("ref_create failed")

    def __del__(self):
        try:
            if self._h:
                self._lib.ref_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def mean_field(self, dict_text, pos, quat, seed, field, two_d):
        """SolidCloud::calcMeanField over one substitute shape: (mean[3], volume)."""
        arr = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        pos, quat, field, out = arr(pos), arr(quat), arr(field), np.zeros(4)
        P = lambda a: C.c_void_p(a.ctypes.data)
        rc = self._lib.ref_mean_field(self._h, dict_text.encode(), P(pos), P(quat), int(seed), P(field), int(bool(two_d)), P(out))
        if rc != 0:
            raise RuntimeError(f"ref_mean_field failed ({rc})")
        return out[:3].copy(), float(out[3])

    def interact(self, dict_texts, solids, seeds, U, dt, rhof, two_d, solid_range=None, want_lists=True):
        n, nC = len(solids), self.mesh.n_cells
        b, e = (0, n) if solid_range is None else solid_range
        texts = (C.c_char_p * n)(*[t.encode() for t in dict_texts])
        arr = lambda a, dt_: np.ascontiguousarray(a, dtype=dt_)
        pos, quat, vel, om = arr(solids["pos"], np.float64), arr(solids["quat"], np.float64), arr(solids["vel"], np.float64), arr(solids["omega"], np.float64)
        seeds, U = arr(seeds, np.int32), arr(U, np.float64)
        cap = 8 * nC + 64
        out = dict(As=np.zeros(nC), Fs=np.zeros((nC, 3)), Ts=np.zeros(nC), Ct=np.zeros(nC), FT=np.zeros((n, 6)),
                   list_off=np.zeros(3 * n + 1, dtype=np.int32), list_cells=np.zeros(cap if want_lists else 1, dtype=np.int32))
        ms = C.c_double(0.0)
        P = lambda a: C.c_void_p(a.ctypes.data)
        rc = self._lib.ref_interact_full(self._h, n, texts, P(pos), P(quat), P(vel), P(om), P(seeds), P(U), C.c_double(dt), C.c_double(rhof),
                                         int(bool(two_d)), int(b), int(e), P(out["list_off"]), P(out["list_cells"]) if want_lists else None,
                                         C.c_int64(cap), P(out["As"]), P(out["Fs"]), P(out["Ts"]), P(out["Ct"]), P(out["FT"]), C.byref(ms))
        if rc < 0:
            raise RuntimeError(f"ref_interact_full failed ({rc})")
        out["list_cells"] = out["list_cells"][:rc].copy() if want_lists else None
        out["pairs"] = int(rc)
        out["timing_ms"] = float(ms.value)
        return out


def plugin_dict_text(entry) -> str:
    """solidDict entries of one motion / forcer from a Python dict (type=..., keys as the reference's constructors read them)."""
    if not entry:
        return ""
    s = ""
    for key, v in entry.items():
        if isinstance(v, (tuple, list, np.ndarray)):
            s += f"{key} ({' '.join(repr(float(x)) for x in v)});\n"
        elif isinstance(v, str):
            s += f"{key} {v};\n"
        else:
            s += f"{key} {float(v)!r};\n"
    return s


def ref_evolve(shape_texts, motions, forcers, rho, pos, quat, vel, omega, times, dt, n_subiter, gravity, rhof, fluid_ft=None,
               bounds=None, delta=-2.0):
    """SolidCloud::evolve over len(times) steps with the reference's own Solid / motions / forcers (and libcollision when bounds is
    given).  Returns dict(pos, quat, vel, omega, FT, traj[n_steps, n, 13])."""
    lib = C.CDLL(LIB_PATH)
    lib.ref_evolve.restype = C.c_int
    n = len(shape_texts)
    arr = lambda a, shape: np.array(a, dtype=np.float64).reshape(shape).copy()
    enc = lambda texts: (C.c_char_p * n)(*[t.encode() for t in texts])
    pos, quat, vel, omega = arr(pos, (n, 3)), arr(quat, (n, 4)), arr(vel, (n, 3)), arr(omega, (n, 3))
    rho, times, g = arr(rho, (n,)), arr(times, (-1,)), arr(gravity, (3,))
    ft, traj = np.zeros((n, 6)), np.zeros((len(times), n, 13))
    P = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
    fl = None if fluid_ft is None else arr(fluid_ft, (len(times), n, 6))
    bmin = None if bounds is None else arr(bounds[0], (3,))
    bmax = None if bounds is None else arr(bounds[1], (3,))
    rc = lib.ref_evolve(n, enc(shape_texts), enc([plugin_dict_text(m) for m in motions]), enc([plugin_dict_text(f) for f in forcers]),
                        P(rho), P(pos), P(quat), P(vel), P(omega), P(fl), P(g), C.c_double(rhof), len(times), P(times), C.c_double(dt),
                        int(n_subiter), P(bmin), P(bmax), C.c_double(delta), P(ft), P(traj))
    if rc != 0:
        raise RuntimeError(f"ref_evolve failed ({rc})")
    return dict(pos=pos, quat=quat, vel=vel, omega=omega, FT=ft, traj=traj)


def ref_fix_internal(mesh, solids, Ct, U):
    """SolidCloud::fixInternal with the reference's Solid::evalPointVelocity; returns the corrected copy of U."""
    lib = C.CDLL(LIB_PATH)
    arr = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    cc, pos, vel, om, Ct = arr(mesh.cc), arr(solids["pos"]), arr(solids["vel"]), arr(solids["omega"]), arr(Ct)
    U = np.array(U, dtype=np.float64, order="C", copy=True)
    P = lambda a: C.c_void_p(a.ctypes.data)
    rc = lib.ref_fix_internal(int(mesh.n_cells), P(cc), len(solids), P(pos), P(vel), P(om), P(Ct), P(U))
    if rc != 0:
        raise RuntimeError(f"ref_fix_internal failed ({rc})")
    return U


def ref_state_rows(pos, quat, vel, omega, ft, time, two_d) -> str:
    """The cloud.out rows the reference's operator<< / write2D (src/solid.cpp) print for these solids."""
    lib = C.CDLL(LIB_PATH)
    lib.ref_state_rows.restype = C.c_int64
    arr = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    pos, quat, vel, omega, ft = arr(pos), arr(quat), arr(vel), arr(omega), arr(ft)
    buf = C.create_string_buffer(512 * len(pos) + 64)
    P = lambda a: C.c_void_p(a.ctypes.data)
    rc = lib.ref_state_rows(len(pos), P(pos), P(quat), P(vel), P(omega), P(ft), C.c_double(time), int(bool(two_d)), buf, C.c_int64(len(buf)))
    if rc < 0:
        raise RuntimeError("ref_state_rows failed")
    return buf.value.decode()


def ref_shape_props(dict_text, pos=None, quat=None, points=None):
    """Mass properties of a shape from the reference's own constructor and, optionally, Solid::phi01 / Solid::phi at world points."""
    lib = C.CDLL(LIB_PATH)
    props = np.zeros(12)
    P = lambda a: C.c_void_p(a.ctypes.data)
    n = 0 if points is None else len(points)
    arr = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    pos, quat, pts = arr(pos if n else np.zeros(3)), arr(quat if n else [1.0, 0, 0, 0]), arr(points if n else np.zeros((1, 3)))
    inside, phi = np.zeros(max(n, 1), dtype=np.uint8), np.zeros(max(n, 1))
    rc = lib.ref_shape_props(dict_text.encode(), P(props), P(pos), P(quat), n, P(pts), P(inside), P(phi))
    if rc < 0:
        raise RuntimeError(f"ref_shape_props failed ({rc})")
    out = dict(volume=props[0], volumeINV=props[1], radiusB=props[2], moi=props[3:6].copy(), moiINV=props[6:9].copy(), com=props[9:12].copy(),
               finite=bool(rc))
    if n:
        out["inside"], out["phi"] = inside[:n].astype(bool), phi[:n]
    return out

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

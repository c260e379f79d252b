"""Foam-free mesh container: the arrays MeshInfo binds (reference src/meshinfo.h:20-29) as numpy views
over storage built by the C++ helpers in csrc/mesh_host.cpp."""
from __future__ import annotations

import ctypes as C

import numpy as np

import os

from . import capi

_mesh_lib = None


class _MeshLib:
    """libsdfibm_mesh.so: csrc/mesh_host.cpp on its own (no CUDA).  Meshes are built through it, never through the CUDA library."""

    def __init__(self):
        from . import build
        path = build.build_mesh()
        self.lib = C.CDLL(path)
        L = self.lib
        L.sdfibm_mesh_last_error.restype = C.c_char_p
        for name, args in {"sdfibm_mesh_from_polymesh": [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)],
                           "sdfibm_mesh_hex_block": [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_void_p)],
                           "sdfibm_mesh_view": [C.c_void_p, C.POINTER(capi.MeshT)],
                           "sdfibm_mesh_owner_neighbour": [C.c_void_p, C.POINTER(capi.c_int32_p), C.POINTER(capi.c_int32_p)],
                           "sdfibm_mesh_free": [C.c_void_p]}.items():
            fn = getattr(L, name)
            fn.restype = C.c_int
            fn.argtypes = args

    def check(self, rc):
        if rc != 0:
            raise capi.SdfibmError(f"mesh helper error {rc}: {self.lib.sdfibm_mesh_last_error().decode()}")


def _lib():
    global _mesh_lib
    if _mesh_lib is None:
        _mesh_lib = _MeshLib()
    return _mesh_lib


def _view(p, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    ct = C.c_double if dtype == np.float64 else C.c_int32
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(n,))


class Mesh:
    """Owns a sdfibm_mesh_storage; exposes numpy views and the sdfibm_mesh_t view struct."""

    def __init__(self, handle):
        self._h = handle
        ml = _lib()
        lib = ml.lib
        self.view = capi.MeshT()
        ml.check(lib.sdfibm_mesh_view(self._h, C.byref(self.view)))
        v = self.view
        self.n_cells, self.n_points, self.n_faces, self.n_internal = v.n_cells, v.n_points, v.n_faces, v.n_internal_faces
        self.points = _view(v.points, 3 * v.n_points, np.float64).reshape(-1, 3)
        self.cc = _view(v.cell_centres, 3 * v.n_cells, np.float64).reshape(-1, 3)
        self.V = _view(v.cell_volumes, v.n_cells, np.float64)
        self.Cf = _view(v.face_centres, 3 * v.n_faces, np.float64).reshape(-1, 3)
        self.Sf = _view(v.face_areas, 3 * v.n_faces, np.float64).reshape(-1, 3)
        self.cp_off = _view(v.cell_points_off, v.n_cells + 1, np.int32)
        self.cp = _view(v.cell_points, int(self.cp_off[-1]), np.int32)
        self.cf_off = _view(v.cell_faces_off, v.n_cells + 1, np.int32)
        self.cf = _view(v.cell_faces, int(self.cf_off[-1]), np.int32)
        self.fp_off = _view(v.face_points_off, v.n_faces + 1, np.int32)
        self.fp = _view(v.face_points, int(self.fp_off[-1]), np.int32)
        self.nb_off = _view(v.cell_cells_off, v.n_cells + 1, np.int32)
        self.nb = _view(v.cell_cells, int(self.nb_off[-1]), np.int32)
        po, pn = capi.c_int32_p(), capi.c_int32_p()
        ml.check(lib.sdfibm_mesh_owner_neighbour(self._h, C.byref(po), C.byref(pn)))
        self.owner = _view(po, v.n_faces, np.int32)
        self.neighbour = _view(pn, v.n_internal_faces, np.int32)
        self.bounds_min = np.array(list(v.bounds_min))
        self.bounds_max = np.array(list(v.bounds_max))

    def __del__(self):
        try:
            if self._h:
                _lib().lib.sdfibm_mesh_free(self._h)
                self._h = None
        except Exception:
            pass

    @classmethod
    def from_polymesh(cls, points, face_off, face_pts, owner, neighbour):
        ml = _lib()
        lib = ml.lib
        points = np.ascontiguousarray(points, dtype=np.float64)
        face_off = np.ascontiguousarray(face_off, dtype=np.int32)
        face_pts = np.ascontiguousarray(face_pts, dtype=np.int32)
        owner = np.ascontiguousarray(owner, dtype=np.int32)
        neighbour = np.ascontiguousarray(neighbour, dtype=np.int32)
        h = C.c_void_p()
        ml.check(lib.sdfibm_mesh_from_polymesh(
            points.shape[0], capi.ptr(points), len(owner), capi.ptr(face_off), capi.ptr(face_pts),
            capi.ptr(owner), len(neighbour), capi.ptr(neighbour), C.byref(h)))
        return cls(h)

    @classmethod
    def hex_block(cls, n, x0=(0.0, 0.0, 0.0), dx=(1.0, 1.0, 1.0)):
        ml = _lib()
        lib = ml.lib
        h = C.c_void_p()
        a = (C.c_double * 3)(*[float(t) for t in x0])
        b = (C.c_double * 3)(*[float(t) for t in dx])
        ml.check(lib.sdfibm_mesh_hex_block(int(n[0]), int(n[1]), int(n[2]), a, b, C.byref(h)))
        return cls(h)

    @classmethod
    def hex_block_with_points(cls, n, points):
        """blockMesh topology of an n[0] x n[1] x n[2] block with the given (e.g. file-read) point coordinates."""
        t = cls.hex_block(n)
        points = np.ascontiguousarray(points, dtype=np.float64)
        assert points.shape == t.points.shape
        return cls.from_polymesh(points, t.fp_off.copy(), t.fp.copy(), t.owner.copy(), t.neighbour.copy())

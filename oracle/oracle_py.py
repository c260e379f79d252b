"""ctypes wrapper of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY (see oracle/oracle.cpp header).

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
_lib = None
_VP = C.c_void_p


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle`")
        lib = C.CDLL(LIB_PATH)
        lib.oracle_create.restype = _VP
        lib.oracle_create.argtypes = [_VP, C.c_int]
        lib.oracle_destroy.argtypes = [_VP]
        lib.oracle_nearest_cell.restype = C.c_int
        lib.oracle_nearest_cell.argtypes = [_VP, _VP]
        lib.oracle_eval_points.argtypes = [_VP, _VP, _VP, C.c_int64, _VP, _VP]
        lib.oracle_interact.restype = C.c_int
        lib.oracle_interact.argtypes = [_VP, _VP, _VP, C.c_int, C.c_int, C.c_int, _VP, C.c_double, C.c_double, C.c_int,
                                        _VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_int64, _VP]
        lib.oracle_fix_internal.argtypes = [_VP, _VP, _VP, C.c_int, _VP, _VP]
        lib.oracle_collide.restype = C.c_int
        lib.oracle_collide.argtypes = [_VP, _VP, _VP, _VP, C.c_int, C.c_double, _VP, C.c_int64, _VP, _VP]
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Oracle:
    """CPU oracle bound to one mesh (sdfibm_b200.mesh.Mesh)."""

    def __init__(self, mesh, two_d: bool):
        self.mesh = mesh  # keeps the storage alive
        self.two_d = bool(two_d)
        self._h = load().oracle_create(C.addressof(mesh.view), int(two_d))

    def __del__(self):
        try:
            if self._h:
                load().oracle_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def nearest_cell(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        return load().oracle_nearest_cell(self._h, _p(p))

    def interact(self, shapes, solids, U, dt, rhof, faithful=False, want_lists=True, solid_range=None, own_vertex_count=False):
        nC = self.mesh.n_cells
        n = len(solids)
        shapes = np.ascontiguousarray(shapes)
        solids = np.ascontiguousarray(solids)
        U = np.ascontiguousarray(U, dtype=np.float64)
        out = dict(As=np.empty(nC), Fs=np.empty((nC, 3)), Ts=np.empty(nC), Ct=np.empty(nC), FT=np.empty((n, 6)))
        timing = np.zeros(2)
        b, e = (0, n) if solid_range is None else solid_range
        off = np.zeros(3 * n + 1, dtype=np.int32) if want_lists else None
        cells = None
        cap = 0
        if want_lists:
            # first call sizes the lists (cells=None), second fills them
            cap = 1 << 16
        while True:
            if want_lists:
                cells = np.empty(cap, dtype=np.int32)
            rc = load().oracle_interact(self._h, _p(shapes), _p(solids), n, b, e, _p(U), float(dt), float(rhof),
                                        int(faithful) | (2 if own_vertex_count else 0), _p(out["As"]), _p(out["Fs"]), _p(out["Ts"]), _p(out["Ct"]),
                                        _p(out["FT"]), _p(off), _p(cells), cap, _p(timing))
            if rc == 4 and want_lists:
                cap = int(off[-1]) + 16
                continue
            if rc != 0:
                raise RuntimeError(f"oracle_interact failed rc={rc}")
            break
        out["timing_ms"] = timing
        if want_lists:
            out["list_off"] = off
            out["list_cells"] = cells[: off[-1]].copy()
        return out

    def fix_internal(self, shapes, solids, Ct, U):
        U = np.array(U, dtype=np.float64, copy=True)
        shapes = np.ascontiguousarray(shapes)
        solids = np.ascontiguousarray(solids)
        Ct = np.ascontiguousarray(Ct, dtype=np.float64)
        load().oracle_fix_internal(self._h, _p(shapes), _p(solids), len(solids), _p(Ct), _p(U))
        return U

    def collide(self, shapes, solids, delta, force_torque=None):
        shapes = np.ascontiguousarray(shapes)
        solids = np.ascontiguousarray(solids)
        bmin = np.ascontiguousarray(self.mesh.bounds_min, dtype=np.float64)
        bmax = np.ascontiguousarray(self.mesh.bounds_max, dtype=np.float64)
        n = len(solids)
        npairs = C.c_int64(0)
        cap = 1 << 16
        ft0 = np.zeros((n, 6)) if force_torque is None else np.array(force_torque, dtype=np.float64)
        while True:
            pairs = np.empty((cap, 2), dtype=np.int32)
            ft = ft0.copy()
            rc = load().oracle_collide(_p(bmin), _p(bmax), _p(shapes), _p(solids), n, float(delta), _p(pairs), cap,
                                       C.byref(npairs), _p(ft))
            if rc == 4:
                cap = int(npairs.value) + 16
                continue
            break
        return pairs[: npairs.value].copy(), ft


_programs_keepalive = None


def set_programs(ops):
    """The op table the SDFIBM_SHAPE_PROGRAM records of later calls point into (process-wide; kept alive here)."""
    global _programs_keepalive
    _programs_keepalive = None if ops is None else np.ascontiguousarray(ops)
    load().oracle_set_programs(None if ops is None else _p(_programs_keepalive))


def eval_points(shapes, solid, pts):
    shapes = np.ascontiguousarray(shapes)
    solid = np.ascontiguousarray(solid)
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    n = pts.shape[0]
    inside = np.empty(n, dtype=np.int32)
    phi = np.empty(n)
    load().oracle_eval_points(_p(shapes), _p(solid), _p(pts), n, _p(inside), _p(phi))
    return inside.astype(bool), phi

"""ctypes binding of include/sdfibm_b200_host.h: the C++ host façade (sdfibm::SolidCloud) from Python.

`HostCloud` mirrors the calls main.cpp makes on SolidCloud (reference src/main.cpp:38-39,66,82-83,87,101).  The C++
library does the solidDict parsing, plugin construction, rigid-body integration and file output; the coupling itself
runs in libsdfibm_b200.so.  `write_solid_dict` renders a solidDict (the reference's schema) from Python data.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsdfibm_host.so")
RUNNER_PATH = os.path.join(HERE, "sdfibm_b200_run")
_VP = C.c_void_p

SYMBOLS = {
    "sdfibm_host_last_error": (C.c_char_p, []),
    "sdfibm_host_create": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(capi.MeshT), C.c_double, C.c_double, _VP, C.POINTER(_VP)]),
    "sdfibm_host_destroy": (C.c_int, [_VP]),
    "sdfibm_host_field": (C.c_int, [_VP, C.c_char_p, C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_int64)]),
    "sdfibm_host_is_on_fluid": (C.c_int, [_VP, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "sdfibm_host_interact": (C.c_int, [_VP, C.c_double, C.c_double]),
    "sdfibm_host_evolve": (C.c_int, [_VP, C.c_double, C.c_double]),
    "sdfibm_host_save_state": (C.c_int, [_VP]),
    "sdfibm_host_fix_internal": (C.c_int, [_VP, C.c_double]),
    "sdfibm_host_save_restart": (C.c_int, [_VP, C.c_char_p]),
    "sdfibm_host_n_solids": (C.c_int, [_VP, C.POINTER(C.c_int)]),
    "sdfibm_host_get_solids": (C.c_int, [_VP, _VP]),
    "sdfibm_host_get_forces": (C.c_int, [_VP, _VP, _VP]),
    "sdfibm_host_get_masses": (C.c_int, [_VP, _VP]),
    "sdfibm_host_mean_field": (C.c_int, [_VP, _VP]),
    "sdfibm_host_write_vof": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(capi.MeshT), C.c_char_p, _VP, C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "sdfibm_host_set_collision_delta": (C.c_int, [_VP, C.c_double]),
    "sdfibm_host_reset_subiterations": (C.c_int, []),
    "sdfibm_host_factory_has": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_int)]),
    "sdfibm_host_register_untagged_shape": (C.c_int, []),
    "sdfibm_host_shape_record": (C.c_int, [C.c_char_p, C.c_char_p, _VP, C.POINTER(C.c_double)]),
    "sdfibm_host_shape_eval": (C.c_int, [C.c_char_p, C.c_char_p, _VP, _VP, _VP, C.c_int64, _VP, _VP]),
}

_lib = None


class HostError(RuntimeError):
    pass


def use_library(path=None):
    """Bind this module to another build of the host façade that exports the same C entry points (tests: the façade compiled
    with the reference's own plugin headers instead of the built-in ones).  None = back to the in-tree libsdfibm_host.so."""
    global _lib, LIB_PATH
    _lib = None
    LIB_PATH = path or os.path.join(HERE, "libsdfibm_host.so")


def load():
    global _lib
    if _lib is not None:
        return _lib
    capi.load()  # the CUDA library first (same directory; the host library links against it)
    if not os.path.exists(LIB_PATH):
        raise HostError(f"{LIB_PATH} is missing: build it with `python -m sdfibm_b200.build`")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise HostError(load().sdfibm_host_last_error().decode())


def factory_has(kind: str, type_name: str) -> bool:
    f = C.c_int(0)
    check(load().sdfibm_host_factory_has(kind.encode(), type_name.encode(), C.byref(f)))
    return bool(f.value)


def shape_record(dictfile: str, shape_name: str):
    rec = np.zeros((), dtype=capi.SHAPE_DTYPE)
    props = (C.c_double * 6)()
    check(load().sdfibm_host_shape_record(dictfile.encode(), shape_name.encode(), C.c_void_p(rec.ctypes.data), props))
    return rec, dict(volume=props[0], volumeINV=props[1], radiusB=props[2], moi=(props[3], props[4], props[5]))


def shape_eval(dictfile: str, shape_name: str, pos, quat, points):
    points = np.ascontiguousarray(points, dtype=np.float64)
    n = len(points)
    inside = np.zeros(n, dtype=np.int32)
    phi = np.zeros(n)
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    quat = np.ascontiguousarray(quat, dtype=np.float64)
    check(load().sdfibm_host_shape_eval(dictfile.encode(), shape_name.encode(), capi.ptr(pos), capi.ptr(quat), capi.ptr(points), n,
                                        capi.ptr(inside), capi.ptr(phi)))
    return inside.astype(bool), phi


def write_vof(dictfile: str, case_dir: str, mesh, field_name: str = "alpha.water"):
    """tool_vof: VofCloud(dictfile, mesh).writeVOF(field_name).  Returns (alpha[n_cells], total volume, n_solids, n_planes);
    the field is also written to <case_dir>/0_<field_name>."""
    alpha = np.empty(mesh.n_cells)
    tot = C.c_double()
    ns, npl = C.c_int(), C.c_int()
    check(load().sdfibm_host_write_vof(dictfile.encode(), case_dir.encode(), C.byref(mesh.view), field_name.encode(), capi.ptr(alpha),
                                       C.byref(tot), C.byref(ns), C.byref(npl)))
    return alpha, float(tot.value), int(ns.value), int(npl.value)


def write_vof_dict(path, on_twod, shapes, solids, planes=()):
    """A solidDict of tool_vof's flavour (tool_vof/example/solidDict): meta.on_twod, shapes{}, solids{ shp_name pos euler }, planes{}."""
    def body(title, items, prefix):
        out = [f"{title}\n{{"]
        for i, b in enumerate(items):
            out.append(f"    {prefix}{i}\n    {{")
            out.append(f"        shp_name {b['shp_name']};")
            out.append(f"        pos {_fmt(tuple(b['pos']))};")
            out.append(f"        euler {_fmt(tuple(b.get('euler', (0, 0, 0))))};")
            out.append("    }")
        out.append("}")
        return "\n".join(out)
    lines = ["meta\n{", f"    on_twod {1 if on_twod else 0};", "}", "shapes\n{"]
    for name, d in shapes.items():
        lines.append(f"    {name}\n    {{")
        lines.append(f"        name {name};")
        for k, v in d.items():
            lines.append(f"        {k} {_fmt(v)};")
        lines.append("    }")
    lines.append("}")
    lines.append(body("solids", solids, "solid"))
    if planes:
        lines.append(body("planes", planes, "plane"))
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return path


class HostCloud:
    """sdfibm::SolidCloud on a Foam-free mesh."""

    def __init__(self, dictfile: str, case_dir: str, mesh, rho_fluid: float = 1.0, start_time: float = 0.0, U_init=None):
        self._lib = load()
        self._h = _VP()
        self.mesh = mesh
        u = None if U_init is None else np.ascontiguousarray(U_init, dtype=np.float64)
        check(self._lib.sdfibm_host_create(dictfile.encode(), case_dir.encode(), C.byref(mesh.view), float(rho_fluid), float(start_time),
                                           capi.ptr(u), C.byref(self._h)))

    def close(self):
        if self._h:
            self._lib.sdfibm_host_destroy(self._h)
            self._h = _VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def field(self, name: str) -> np.ndarray:
        p = C.POINTER(C.c_double)()
        n = C.c_int64(0)
        check(self._lib.sdfibm_host_field(self._h, name.encode(), C.byref(p), C.byref(n)))
        a = np.ctypeslib.as_array(p, shape=(n.value,))
        return a.reshape(-1, 3) if name in ("U", "Fs") else a

    def flags(self):
        a, b = C.c_int(0), C.c_int(0)
        check(self._lib.sdfibm_host_is_on_fluid(self._h, C.byref(a), C.byref(b)))
        return bool(a.value), bool(b.value)

    def interact(self, t, dt): check(self._lib.sdfibm_host_interact(self._h, float(t), float(dt)))
    def evolve(self, t, dt): check(self._lib.sdfibm_host_evolve(self._h, float(t), float(dt)))
    def save_state(self): check(self._lib.sdfibm_host_save_state(self._h))
    def fix_internal(self, dt): check(self._lib.sdfibm_host_fix_internal(self._h, float(dt)))
    def save_restart(self, path): check(self._lib.sdfibm_host_save_restart(self._h, path.encode()))
    def set_collision_delta(self, d): check(self._lib.sdfibm_host_set_collision_delta(self._h, float(d)))

    @property
    def n_solids(self):
        n = C.c_int(0)
        check(self._lib.sdfibm_host_n_solids(self._h, C.byref(n)))
        return n.value

    def solids(self) -> np.ndarray:
        out = np.zeros(self.n_solids, dtype=capi.SOLID_DTYPE)
        check(self._lib.sdfibm_host_get_solids(self._h, capi.ptr(out)))
        return out

    def forces(self):
        n = self.n_solids
        ft, fl = np.zeros((n, 6)), np.zeros((n, 6))
        check(self._lib.sdfibm_host_get_forces(self._h, capi.ptr(ft), capi.ptr(fl)))
        return ft, fl

    def mean_field(self):
        m = np.zeros((self.n_solids, 3))
        check(self._lib.sdfibm_host_mean_field(self._h, capi.ptr(m)))
        return m

    def masses(self):
        m = np.zeros(self.n_solids)
        check(self._lib.sdfibm_host_get_masses(self._h, capi.ptr(m)))
        return m


# ---- solidDict writer (the reference's schema) -------------------------------------
def _fmt(v):
    if isinstance(v, (tuple, list, np.ndarray)):
        return "(" + " ".join(repr(float(x)) for x in v) + ")"
    if isinstance(v, bool):
        return "1" if v else "0"
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    if isinstance(v, (float, np.floating)):
        return repr(float(v))
    return str(v)


def write_solid_dict(path, meta, shapes, motions, materials, solids, forces=None):
    """shapes / motions / materials / forces: {name: dict(type=..., key=value...)}; solids: list of dicts."""
    def block(title, entries):
        out = [title, "{"]
        for name, d in entries.items():
            out += [f"    {name}", "    {", f"        name {name};"]
            out += [f"        {k} {_fmt(v)};" for k, v in d.items()]
            out += ["    }"]
        return out + ["}", ""]

    lines = ["FoamFile", "{", "    version 2.0;", "    format ascii;", "    class dictionary;", "    object solidDict;", "}", "",
             "// written by sdfibm_b200.hostapi.write_solid_dict", "meta", "{"]
    lines += [f"    {k} {_fmt(v)};" for k, v in meta.items()] + ["}", ""]
    lines += block("shapes", shapes)
    if forces:
        lines += block("forces", forces)
    lines += block("motions", motions) + block("materials", materials)
    lines += ["solids", "{"]
    for i, s in enumerate(solids):
        lines += [f"    solid{i}", "    {"] + [f"        {k} {_fmt(v)};" for k, v in s.items()] + ["    }"]
    lines += ["}", ""]
    with open(path, "w") as f:
        f.write("\n".join(lines))
    return path

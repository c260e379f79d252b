"""ctypes binding of include/sdfibm_b200.h (the C-ABI drop-in boundary).

The library is loaded from the in-tree build (sdfibm_b200/libsdfibm_b200.so).  There is no CPU
fallback: if the library is missing, loading raises; if no CUDA device is usable, every compute
entry returns an error that is raised as SdfibmError.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDFIBM_B200_LIB") or os.path.join(HERE, "libsdfibm_b200.so")   # the override is for kernel-variant experiments

SHAPE_TAGS = {
    "Plane": 0, "Circle": 1, "Sphere": 2, "Ellipse": 3, "Ellipsoid": 4,
    "Rectangle": 5, "Box": 6, "Circle_Tail": 7, "Circle_TwoTail": 8,
}
SHAPE_PROGRAM_TAG = 9
# sdfibm_sdf_op_t and its opcodes (include/sdfibm_b200.h)
SDF_OP_DTYPE = np.dtype([("op", "<i4"), ("pad_", "<i4"), ("a", "<f8", 4)], align=True)
SDF_OPS = {"POINT": 0, "POINT_2D": 1, "OFFSET": 2, "ROT30": 3, "ROT45": 4, "ROT60": 5, "ROT90": 6, "ROTTH": 7, "FLIPX": 8, "FLIPY": 9,
           "CIRCLE": 16, "RECTANGLE": 17, "BOX": 18, "ELLIPSE": 19, "ELLIPSOID": 20, "HALFSPACE": 21,
           "UNION": 32, "INTERSECT": 33, "DIFF": 34}

# numpy dtypes that mirror the POD records (sdfibm_shape_t, sdfibm_solid_t)
SHAPE_DTYPE = np.dtype(
    [("tag", "<i4"), ("finite", "<i4"), ("radiusB", "<f8"), ("com", "<f8", 3), ("p", "<f8", 8)], align=True
)
SOLID_DTYPE = np.dtype(
    [("pos", "<f8", 3), ("quat", "<f8", 4), ("vel", "<f8", 3), ("omega", "<f8", 3), ("shape", "<i4"), ("pad_", "<i4")],
    align=True,
)
assert SHAPE_DTYPE.itemsize == 104 and SOLID_DTYPE.itemsize == 112 and SDF_OP_DTYPE.itemsize == 40

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class MeshT(C.Structure):
    _fields_ = [
        ("n_cells", C.c_int32), ("n_points", C.c_int32), ("n_faces", C.c_int32), ("n_internal_faces", C.c_int32),
        ("points", c_double_p), ("cell_centres", c_double_p), ("cell_volumes", c_double_p),
        ("face_centres", c_double_p), ("face_areas", c_double_p),
        ("cell_points_off", c_int32_p), ("cell_points", c_int32_p),
        ("cell_faces_off", c_int32_p), ("cell_faces", c_int32_p),
        ("face_points_off", c_int32_p), ("face_points", c_int32_p),
        ("cell_cells_off", c_int32_p), ("cell_cells", c_int32_p),
        ("bounds_min", C.c_double * 3), ("bounds_max", C.c_double * 3),
    ]


class SdfibmError(RuntimeError):
    pass


# every symbol include/sdfibm_b200.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
SYMBOLS = {
    "sdfibm_version": (C.c_int, []),
    "sdfibm_last_error": (C.c_char_p, []),
    "sdfibm_device_count": (C.c_int, [c_int32_p]),
    "sdfibm_create": (C.c_int, [C.c_int, C.POINTER(_VP)]),
    "sdfibm_destroy": (C.c_int, [_VP]),
    "sdfibm_set_cell_slots": (C.c_int, [_VP, C.c_int]),
    "sdfibm_set_mesh": (C.c_int, [_VP, C.POINTER(MeshT), C.c_int]),
    "sdfibm_set_shapes": (C.c_int, [_VP, _VP, C.c_int]),
    "sdfibm_set_shape_programs": (C.c_int, [_VP, _VP, C.c_int]),
    "sdfibm_interact": (C.c_int, [_VP, _VP, C.c_int, _VP, C.c_double, C.c_double, _VP, _VP, _VP, _VP, _VP]),
    "sdfibm_interact_device": (C.c_int, [_VP, _VP, C.c_int, _VP, C.c_double, C.c_double, _VP, _VP, _VP, _VP, _VP]),
    "sdfibm_interact_device_solids": (C.c_int, [_VP, _VP, C.c_int, C.c_int, _VP, C.c_double, C.c_double, _VP, _VP, _VP, _VP, _VP]),
    "sdfibm_fix_internal": (C.c_int, [_VP, _VP, C.c_int, _VP]),
    "sdfibm_fix_internal_device": (C.c_int, [_VP, _VP, C.c_int, _VP, _VP]),
    "sdfibm_volume_fraction": (C.c_int, [_VP, _VP, C.c_int, _VP, c_double_p]),
    "sdfibm_mean_field": (C.c_int, [_VP, _VP, C.c_int, _VP, _VP, _VP]),
    "sdfibm_mean_field_sums": (C.c_int, [_VP, _VP, C.c_int, _VP, _VP, _VP]),
    "sdfibm_allow_order_free": (C.c_int, [_VP, C.c_int]),
    "sdfibm_apply_forcing_device": (C.c_int, [_VP, _VP, _VP, C.c_double]),
    "sdfibm_download": (C.c_int, [_VP, _VP, _VP, C.c_size_t]),
    "sdfibm_touched_cells": (C.c_int, [_VP, C.c_int64, c_int64_p, _VP, _VP, _VP, _VP, _VP]),
    "sdfibm_comm_unique_id": (C.c_int, [_VP]),
    "sdfibm_comm_init": (C.c_int, [_VP, _VP, C.c_int, C.c_int]),
    "sdfibm_comm_options": (C.c_int, [_VP, C.c_int, C.c_int]),
    "sdfibm_comm_destroy": (C.c_int, [_VP]),
    "sdfibm_allreduce_force_torque": (C.c_int, [_VP, _VP, C.c_int]),
    "sdfibm_comm_last_ms": (C.c_int, [_VP, c_double_p]),
    "sdfibm_candidate_counts": (C.c_int, [_VP, c_int64_p]),
    "sdfibm_candidate_lists": (C.c_int, [_VP, _VP, _VP, C.c_int64]),
    "sdfibm_last_stats": (C.c_int, [_VP, c_int64_p]),
    "sdfibm_alloc_pinned": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "sdfibm_free_pinned": (C.c_int, [_VP]),
    "sdfibm_last_timings": (C.c_int, [_VP, c_double_p]),
    "sdfibm_last_host_timings": (C.c_int, [_VP, c_double_p]),
    "sdfibm_last_aux_timings": (C.c_int, [_VP, c_double_p]),
    "sdfibm_collide": (C.c_int, [_VP, _VP, C.c_int, C.c_double, _VP, C.c_int64, c_int64_p, _VP]),
    "sdfibm_stream": (C.c_int, [_VP, C.POINTER(_VP)]),
    "sdfibm_synchronize": (C.c_int, [_VP]),
    "sdfibm_mesh_from_polymesh": (C.c_int, [C.c_int32, _VP, C.c_int32, _VP, _VP, _VP, C.c_int32, _VP, C.POINTER(_VP)]),
    "sdfibm_mesh_hex_block": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, c_double_p, c_double_p, C.POINTER(_VP)]),
    "sdfibm_mesh_view": (C.c_int, [_VP, C.POINTER(MeshT)]),
    "sdfibm_mesh_owner_neighbour": (C.c_int, [_VP, C.POINTER(c_int32_p), C.POINTER(c_int32_p)]),
    "sdfibm_mesh_free": (C.c_int, [_VP]),
}

_lib = None


def load():
    """Load libsdfibm_b200.so (building is the caller's job: `python -m sdfibm_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SdfibmError(
            f"{LIB_PATH} is missing: build it with `python -m sdfibm_b200.build` (there is no CPU fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().sdfibm_last_error()
        raise SdfibmError(f"sdfibm error {rc}: {msg.decode() if msg else ''}")


def ptr(a):
    """void* of a numpy array (must be C-contiguous) or a raw integer device pointer."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def pinned_like(a):
    """Copy of `a` in page-locked host memory from sdfibm_alloc_pinned (never freed: meant for long-lived per-step buffers)."""
    import numpy as np

    a = np.ascontiguousarray(a)
    p = C.c_void_p()
    check(load().sdfibm_alloc_pinned(max(a.nbytes, 1), C.byref(p)))
    buf = (C.c_char * max(a.nbytes, 1)).from_address(p.value)
    out = np.frombuffer(buf, dtype=a.dtype, count=a.size).reshape(a.shape)
    out[...] = a
    return out

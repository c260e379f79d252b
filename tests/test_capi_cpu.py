"""CPU-side checks of the C-ABI boundary: the library loads, exports every declared symbol, and fails
loudly (no CPU fallback) when no CUDA device is usable."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from sdfibm_b200 import capi, cases
from sdfibm_b200.mesh import Mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, "include", "sdfibm_b200.h")).read()
    declared = set(re.findall(r"\b(sdfibm_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"sdfibm_context", "sdfibm_mesh_storage"}
    lib = capi.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/sdfibm_b200.h but not exported"
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    assert lib.sdfibm_version() >= 100


def test_pod_layouts_match_header():
    assert capi.SHAPE_DTYPE.itemsize == 104
    assert capi.SOLID_DTYPE.itemsize == 112
    assert C.sizeof(capi.MeshT) == 168


def test_no_cpu_fallback_without_device():
    lib = capi.load()
    n = C.c_int32(0)
    rc = lib.sdfibm_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = lib.sdfibm_create(0, C.byref(h))
    assert rc != 0
    assert b"no CUDA device" in lib.sdfibm_last_error()


def test_hex_block_geometry():
    m = Mesh.hex_block((4, 3, 2), (1.0, 2.0, 3.0), (0.5, 0.25, 2.0))
    assert (m.n_cells, m.n_points) == (24, 5 * 4 * 3)
    assert np.allclose(m.V, 0.5 * 0.25 * 2.0)
    assert np.allclose(m.cc[0], (1.25, 2.125, 4.0))
    # closed cells: sum of outward face area vectors vanishes
    for c in range(m.n_cells):
        s = np.zeros(3)
        for f in m.cf[m.cf_off[c]:m.cf_off[c + 1]]:
            s += m.Sf[f] if m.owner[f] == c else -m.Sf[f]
        assert np.abs(s).max() < 1e-14
    assert np.array_equal(m.bounds_min, (1.0, 2.0, 3.0)) and np.array_equal(m.bounds_max, (3.0, 2.75, 7.0))


def test_prism_mesh_is_closed_and_positive():
    m = cases.prism_mesh(5, 4)
    assert m.n_cells == 40 and (np.diff(m.cp_off) == 6).all() and (np.diff(m.cf_off) == 5).all()
    assert np.allclose(m.V, 0.5) and np.isclose(m.V.sum(), 20.0)


def test_shape_records():
    from sdfibm_b200.shapes import make_shape, quat_from_euler_xyz_deg
    s = make_shape("Circle_TwoTail", radius=0.3, ratio=1, thickness=0.1)
    assert s["p"][2] == (1 + 1) * 0.5 * 0.3 and s["p"][3] == 0.05 and s["radiusB"] == 2 * s["p"][2]
    s = make_shape("Plane")
    assert s["finite"] == 0 and s["tag"] == 0
    with pytest.raises(ValueError):
        make_shape("Torus", radius=1)
    q = quat_from_euler_xyz_deg((0, 0, 90))
    assert np.allclose(q, (np.sqrt(0.5), 0, 0, np.sqrt(0.5)))

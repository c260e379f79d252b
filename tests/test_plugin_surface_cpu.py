"""Drop-in proof of the plugin surface: the REFERENCE's own plugin headers (libmotion/motion*.h, libforcer/{constant,spring,
magnetic}.h, libshape/{sphere,ellipsoid}.h + sdf/sdf.h), compiled unmodified from where they lie against this repository's
interface headers and linked into a copy of the host façade INSTEAD of the built-in plugin sets (tools/build_plugin_proof.py ->
build/plugin_proof/libplugin_proof.so, git-ignored), drive sdfibm::SolidCloud to the same results as the built-in plugins:
evolve (all seven motions, all three forcers, 20 sub-iterations) bit for bit on the CPU; interact / fixInternal on the GPU."""
import os
import sys

import numpy as np
import pytest

import host_cases as hc
from sdfibm_b200 import hostapi
from sdfibm_b200.mesh import Mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import build_plugin_proof  # noqa: E402


def _proof_lib():
    lib = build_plugin_proof.build() if os.path.isdir("/root/reference/src") else (build_plugin_proof.LIB if os.path.exists(build_plugin_proof.LIB) else None)
    if not lib:
        pytest.skip("needs the reference tree (or the prebuilt build/plugin_proof/libplugin_proof.so)")
    return lib


SOLIDS = [
    dict(shp_name="sph", mot_name="free", mat_name="heavy", for_name="push", pos=(0.0, 0.0, 0.0), vel=(0.1, 0.2, -0.1), omega=(0.3, -0.2, 0.5)),
    dict(shp_name="elo", mot_name="free", mat_name="light", for_name="spring", pos=(1.0, 0.5, -0.5), euler=(30, -20, 45), vel=(0, 0.1, 0), omega=(0.1, 0.2, 0.3)),
    dict(shp_name="sph", mot_name="mask", mat_name="heavy", pos=(-1.0, 1.0, 0.0), vel=(1, 1, 1), omega=(1, 1, 1)),
    dict(shp_name="elo", mot_name="spin", mat_name="light", pos=(-2.0, 0.0, 1.0), vel=(1, 1, 1)),
    dict(shp_name="sph", mot_name="spinfree", mat_name="heavy", for_name="mag", pos=(-2.0, 1.0, 1.0), vel=(0.3, -0.1, 0.2), euler=(0, 0, 15)),
    dict(shp_name="elo", mot_name="const", mat_name="light", pos=(2.0, -1.0, 0.0)),
    dict(shp_name="sph", mot_name="sine", mat_name="heavy", pos=(0.0, -1.5, 0.5)),
    dict(shp_name="sph", mot_name="rotor", mat_name="light", pos=(0.7, 0.0, 0.0)),
    dict(shp_name="elo", mot_name="gate", mat_name="light", pos=(3.0, 0.0, 0.0)),
]
SHAPES = {k: hc.SHAPES[k] for k in ("sph", "elo")}       # the two shape headers linked from the reference
META = dict(on_fluid=0, on_twod=0, gravity=(0.0, -9.8, 0.0))


def _trajectory(tmp, steps=6):
    path = hc.write_case(tmp, META, SOLIDS, shapes=SHAPES)
    hostapi.load().sdfibm_host_reset_subiterations()
    mesh = Mesh.hex_block((4, 4, 4), (-2, -2, -2), (1.0, 1.0, 1.0))
    cloud = hostapi.HostCloud(path, str(tmp), mesh, rho_fluid=1.0, start_time=0.5)   # restart mode: no device call
    out = []
    t, dt = 0.5, 0.01
    for _ in range(steps):
        t += dt
        cloud.evolve(t, dt)
        cloud.save_state()
        s = cloud.solids()
        out.append((s["pos"].copy(), s["quat"].copy(), s["vel"].copy(), s["omega"].copy(), cloud.forces()[0].copy(), np.array(cloud.masses())))
    cloud.close()
    return out, open(os.path.join(str(tmp), "cloud.out")).read()


def test_reference_plugin_headers_compile_against_the_facade_and_evolve_identically(tmp_path):
    lib = _proof_lib()
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    try:
        mine, rows_mine = _trajectory(tmp_path / "a")
        hostapi.use_library(lib)
        for kind, names in (("motion", ["Motion000002", "Motion110002", "Motion222000", "MotionSineDirectional", "Motion01Mask", "MotionRotor", "MotionOpenClose"]),
                            ("forcer", ["Constant", "Spring", "Magnetic"]), ("shape", ["Sphere", "Ellipsoid"])):
            for n in names:
                assert hostapi.factory_has(kind, n), (kind, n)
        assert not hostapi.factory_has("shape", "Circle")      # only what was linked from the reference is registered
        theirs, rows_theirs = _trajectory(tmp_path / "b")
    finally:
        hostapi.use_library(None)
    assert len(mine) == len(theirs) == 6
    for a, b in zip(mine, theirs):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    assert rows_mine == rows_theirs                             # cloud.out, character for character


@pytest.mark.gpu
def test_reference_plugins_interact_on_the_gpu_like_the_builtin_ones(tmp_path):
    lib = _proof_lib()
    meta = dict(on_fluid=1, on_twod=0, gravity=(0.0, 0.0, 0.0))
    solids = [dict(shp_name="sph", mot_name="free", mat_name="heavy", pos=(1.3, 1.1, 0.9), vel=(0.1, 0.0, -0.1), omega=(0.0, 0.2, 0.1)),
              dict(shp_name="elo", mot_name="free", mat_name="light", pos=(-0.9, -0.7, 0.4), euler=(30, -20, 45))]
    mesh = Mesh.hex_block((24, 24, 24), (-3, -3, -3), (0.25, 0.25, 0.25))
    U = np.random.RandomState(2).standard_normal((mesh.n_cells, 3))
    res = []
    try:
        for which, sub in ((None, "a"), (lib, "b")):
            d = tmp_path / sub
            d.mkdir()
            hostapi.use_library(which)
            path = hc.write_case(d, meta, solids, shapes=SHAPES)
            cloud = hostapi.HostCloud(path, str(d), mesh, rho_fluid=1.2, start_time=1.0, U_init=U)
            cloud.interact(1.0, 1e-3)
            cloud.evolve(1.0, 1e-3)
            cloud.fix_internal(1e-3)
            res.append({k: cloud.field(k).copy() for k in ("As", "Fs", "Ts", "Ct", "U")} | {"solids": cloud.solids().copy()})
            cloud.close()
    finally:
        hostapi.use_library(None)
    for k in ("As", "Fs", "Ts", "Ct"):
        assert np.array_equal(res[0][k], res[1][k]), k
    # the per-solid force sums are atomic accumulations (not bit-reproducible run to run): the solid states after evolve, and
    # the velocities fixInternal writes from them, agree to rounding
    assert np.abs(res[0]["U"] - res[1]["U"]).max() <= 1e-12
    for k in ("pos", "quat", "vel", "omega"):
        assert np.abs(res[0]["solids"][k] - res[1]["solids"][k]).max() <= 1e-12, k
    assert res[0]["As"].max() == 1.0

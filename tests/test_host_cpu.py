"""Host façade (C++ sdfibm::SolidCloud over the C ABI) — everything that needs no GPU: solidDict parsing, plugin
registries, shape lowering and mass properties, the rigid-body integrator with motions / forcers / gravity against the
oracle restatement, cloud.out and the restart round trip, and the loud failure of the device calls without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import host_cases as hc
from oracle import host_oracle as ho
from oracle.oracle_py import eval_points
from sdfibm_b200 import capi, hostapi
from sdfibm_b200.mesh import Mesh
from sdfibm_b200.shapes import make_solids, quat_from_euler_xyz_deg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
META_DEM_3D = dict(on_fluid=0, on_twod=0, gravity=(0.0, -9.81, 0.5))


def has_gpu():
    n = C.c_int32(0)
    return capi.load().sdfibm_device_count(C.byref(n)) == 0 and n.value > 0


def test_host_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "sdfibm_b200_host.h")).read()
    declared = set(re.findall(r"\b(sdfibm_host_[a-z_0-9]+)\s*\(", hdr))
    lib = hostapi.load()
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(hostapi.SYMBOLS), declared ^ set(hostapi.SYMBOLS)


def test_builtin_plugins_are_registered():
    for t in ("Circle", "Sphere", "Ellipse", "Ellipsoid", "Rectangle", "Box", "Circle_Tail", "Circle_TwoTail", "Plane"):
        assert hostapi.factory_has("shape", t), t
    for t in ("Motion01Mask", "Motion000002", "Motion110002", "Motion222000", "MotionSineDirectional", "MotionRotor", "MotionOpenClose"):
        assert hostapi.factory_has("motion", t), t
    for t in ("Constant", "Spring", "Magnetic"):
        assert hostapi.factory_has("forcer", t), t
    assert not hostapi.factory_has("shape", "Torus")
    assert not hostapi.factory_has("motion", "Motion999999")
    assert not hostapi.factory_has("forcer", "Gravity")


def test_shape_lowering_and_mass_properties(tmp_path):
    path = hc.write_case(tmp_path, META_DEM_3D, [])
    for name in hc.SHAPES:
        rec, props = hostapi.shape_record(path, name)
        ref = hc.shape_record(name)
        for f in ("tag", "finite", "radiusB"):
            assert rec[f] == ref[f], (name, f)
        assert np.array_equal(rec["com"], ref["com"]) and np.array_equal(rec["p"], ref["p"]), name
        vol, vinv, minv = hc.mass_props(name)
        assert props["volume"] == vol and props["volumeINV"] == vinv, name
        if name != "plane":
            assert np.allclose(1.0 / np.array(props["moi"]), minv, rtol=1e-15)


def test_host_shape_evaluation_matches_oracle(tmp_path):
    path = hc.write_case(tmp_path, META_DEM_3D, [])
    rng = np.random.RandomState(3)
    pts = rng.uniform(-0.8, 0.8, size=(4000, 3))
    pos = np.array([0.05, -0.1, 0.02])
    quat = np.array(quat_from_euler_xyz_deg((20, -35, 50)))
    for name in hc.SHAPES:
        inside, phi = hostapi.shape_eval(path, name, pos, quat, pts)
        S = make_solids(1)
        S[0]["pos"], S[0]["quat"] = pos, quat
        ri, rp = eval_points(np.array([hc.shape_record(name)]), S[0], pts)
        assert np.array_equal(inside, ri), name
        assert np.array_equal(phi, rp), name


def _solids_3d():
    return [
        dict(shp_name="sph", mot_name="free", mat_name="heavy", pos=(0.2, 1.5, 0.1), vel=(0.1, 0.0, -0.1), euler=(10, 20, 30), omega=(0.3, -0.2, 0.5)),
        dict(shp_name="elo", mot_name="free", mat_name="light", for_name="spring", pos=(1.0, 0.7, -0.3), euler=(0, 45, 10), omega=(0, 0.4, 0.1)),
        dict(shp_name="box", mot_name="mask", mat_name="heavy", for_name="push", pos=(-0.5, 0.2, 0.4), vel=(0.2, 0.1, 0.3), omega=(0.1, 0.2, 0.3)),
        dict(shp_name="sph", mot_name="rotor", mat_name="light", pos=(0.7, 0.0, 0.0)),
        dict(shp_name="elo", mot_name="sine", mat_name="heavy", for_name="mag", pos=(0.0, -1.0, 0.5), euler=(30, 0, 0)),
        dict(shp_name="box", mot_name="const", mat_name="light", pos=(2.0, 2.0, 2.0)),
        dict(shp_name="sph", mot_name="spin", mat_name="light", pos=(-2.0, 0.0, 1.0), vel=(1, 1, 1)),
        dict(shp_name="sph", mot_name="spinfree", mat_name="heavy", for_name="mag", pos=(-2.0, 1.0, 1.0), vel=(0.3, -0.1, 0.2), euler=(0, 0, 15)),
        dict(shp_name="box", mot_name="gate", mat_name="light", pos=(3.0, 0.0, 0.0)),
    ]


def test_evolve_matches_oracle_dem_3d(tmp_path):
    """20 sub-iterations per step, all motion and forcer types, gravity with rhof = 0 (on_fluid 0) — restart mode so no
    device call is needed (start_time > 0 skips initialCorrect, reference src/solidcloud.cpp:254-258)."""
    solids = _solids_3d()
    path = hc.write_case(tmp_path, META_DEM_3D, solids)
    hostapi.load().sdfibm_host_reset_subiterations()
    mesh = Mesh.hex_block((4, 4, 4), (-2, -2, -2), (1.0, 1.0, 1.0))
    cloud = hostapi.HostCloud(path, str(tmp_path), mesh, rho_fluid=1.0, start_time=0.5)
    assert cloud.flags() == (False, False) and cloud.n_solids == len(solids)
    ref = hc.oracle_solids(solids)
    assert np.allclose(cloud.masses(), [s.mass for s in ref], rtol=1e-15)
    t, dt = 0.5, 0.01
    for step in range(6):
        t += dt
        cloud.evolve(t, dt)
        ho.evolve(ref, t, dt, 20, META_DEM_3D["gravity"], 0.0)
        cloud.save_state()
        got = cloud.solids()
        x, q, v, om = hc.state_arrays(ref)
        for a, b in ((got["pos"], x), (got["quat"], q), (got["vel"], v), (got["omega"], om)):
            assert np.abs(a - b).max() <= 1e-13 * max(1.0, np.abs(b).max()), step
        ft, _ = cloud.forces()
        assert np.allclose(ft, [s.force + s.torque for s in ref], rtol=1e-12, atol=1e-14)
    # cloud.out: one row per solid per saveState, 1 + 18 columns in 3-D (reference README.md:112-121)
    rows = [r.split() for r in open(os.path.join(str(tmp_path), "cloud.out")).read().strip().split("\n")]
    assert len(rows) == 6 * len(solids) and all(len(r) == 19 for r in rows)
    assert float(rows[-1][0]) == pytest.approx(t)
    last = np.array(rows[-len(solids):], dtype=float)
    assert np.allclose(last[:, 1:4], got["pos"], rtol=1e-6, atol=1e-9)


def test_single_solid_uses_one_subiteration_and_2d_output(tmp_path):
    solids = [dict(shp_name="circ", mot_name="free", mat_name="heavy", pos=(0.1, 0.2, 0.0), vel=(0.1, 0, 0), euler=(0, 0, 30), omega=(0, 0, 1.0))]
    meta = dict(on_fluid=0, on_twod=1, gravity=(0.0, -10.0, 0.0), writeFrequency=2)
    path = hc.write_case(tmp_path, meta, solids)
    hostapi.load().sdfibm_host_reset_subiterations()
    mesh = Mesh.hex_block((4, 4, 1), (-1, -1, -0.5), (0.5, 0.5, 1.0))
    cloud = hostapi.HostCloud(path, str(tmp_path), mesh, 1.0, start_time=1.0)
    ref = hc.oracle_solids(solids)
    for step in range(4):
        cloud.evolve(1.0 + step * 1e-3, 1e-3)
        ho.evolve(ref, 1.0 + step * 1e-3, 1e-3, 1, meta["gravity"], 0.0)   # N_SUBITER forced to 1 (src/solidcloud.cpp:524-526)
        cloud.save_state()
    x, q, v, om = hc.state_arrays(ref)
    got = cloud.solids()
    assert np.abs(got["pos"] - x).max() <= 1e-15 and np.abs(got["quat"] - q).max() <= 1e-15
    rows = [r.split() for r in open(os.path.join(str(tmp_path), "cloud.out")).read().strip().split("\n")]
    assert len(rows) == 2 and all(len(r) == 10 for r in rows)      # writeFrequency 2, 1 + 9 columns in 2-D
    hostapi.load().sdfibm_host_reset_subiterations()


def test_restart_round_trip(tmp_path):
    solids = _solids_3d()
    path = hc.write_case(tmp_path, META_DEM_3D, solids)
    hostapi.load().sdfibm_host_reset_subiterations()
    mesh = Mesh.hex_block((4, 4, 4), (-2, -2, -2), (1.0, 1.0, 1.0))
    cloud = hostapi.HostCloud(path, str(tmp_path), mesh, 1.0, start_time=0.5)
    for step in range(3):
        cloud.evolve(0.5 + 0.01 * step, 0.01)
    restart = os.path.join(str(tmp_path), "solidDict.restart")
    cloud.save_restart(restart)
    again = hostapi.HostCloud(restart, str(tmp_path), mesh, 1.0, start_time=0.53)
    a, b = cloud.solids(), again.solids()
    assert np.array_equal(a["pos"], b["pos"]) and np.array_equal(a["vel"], b["vel"]) and np.array_equal(a["omega"], b["omega"])
    # orientation goes through Euler angles in degrees (src/solidcloud.cpp:638): q and -q are the same rotation
    dq = np.minimum(np.abs(a["quat"] - b["quat"]).max(axis=1), np.abs(a["quat"] + b["quat"]).max(axis=1))
    assert dq.max() < 1e-14
    text = open(restart).read()
    assert text.count("FoamFile") == 1 and "shapes" in text and "motions" in text and "materials" in text


def test_bad_dictionaries_are_reported(tmp_path):
    mesh = Mesh.hex_block((2, 2, 2))
    base = dict(shp_name="sph", mot_name="free", mat_name="heavy", pos=(0.0, 0.0, 0.0))
    cases = [
        (dict(base, shp_name="nope"), None, "shape name"),
        (dict(base, mot_name="nope"), None, "motion name"),
        (dict(base, mat_name="nope"), None, "material name"),
        (dict(base, for_name="nope"), None, "force name"),
        (base, {"bad": dict(type="Torus", radius=1.0)}, "unrecognized object"),
    ]
    for i, (solid, extra_shapes, msg) in enumerate(cases):
        d = tmp_path / f"c{i}"
        d.mkdir()
        shapes = dict(hc.SHAPES, **(extra_shapes or {}))
        path = hc.write_case(d, META_DEM_3D, [solid], shapes=shapes)
        with pytest.raises(hostapi.HostError, match=msg):
            hostapi.HostCloud(path, str(d), mesh, 1.0, start_time=1.0)
    # 2-D requires z = 0 (src/solidcloud.cpp:166-170)
    d = tmp_path / "z"
    d.mkdir()
    path = hc.write_case(d, dict(META_DEM_3D, on_twod=1), [dict(base, shp_name="circ", pos=(0, 0, 0.1))])
    with pytest.raises(hostapi.HostError, match="z=0"):
        hostapi.HostCloud(path, str(d), mesh, 1.0, start_time=1.0)


def test_shape_without_device_tag_is_a_hard_error(tmp_path):
    hostapi.check(hostapi.load().sdfibm_host_register_untagged_shape())
    assert hostapi.factory_has("shape", "TestNoDeviceTag")
    shapes = dict(hc.SHAPES, custom=dict(type="TestNoDeviceTag"))
    path = hc.write_case(tmp_path, META_DEM_3D, [dict(shp_name="custom", mot_name="free", mat_name="heavy", pos=(0, 0, 0))], shapes=shapes)
    with pytest.raises(hostapi.HostError, match="no device tag"):
        hostapi.HostCloud(path, str(tmp_path), Mesh.hex_block((2, 2, 2)), 1.0, start_time=1.0)


def test_device_calls_fail_loudly_without_gpu(tmp_path):
    if has_gpu():
        pytest.skip("a CUDA device is present")
    solids = [dict(shp_name="sph", mot_name="free", mat_name="heavy", pos=(0.0, 0.0, 0.0))]
    path = hc.write_case(tmp_path, dict(META_DEM_3D, on_fluid=1), solids)
    mesh = Mesh.hex_block((4, 4, 4), (-1, -1, -1), (0.5, 0.5, 0.5))
    with pytest.raises(hostapi.HostError, match="no CUDA device"):      # t = 0: initialCorrect() runs interact()
        hostapi.HostCloud(path, str(tmp_path), mesh, 1.0, start_time=0.0)
    cloud = hostapi.HostCloud(path, str(tmp_path), mesh, 1.0, start_time=1.0)
    with pytest.raises(hostapi.HostError, match="no CUDA device"):
        cloud.interact(1.0, 1e-3)
    with pytest.raises(hostapi.HostError, match="no CUDA device"):
        cloud.fix_internal(1e-3)
    cloud.set_collision_delta(0.7)                                       # intended-mode collisions need the device too
    with pytest.raises(hostapi.HostError, match="no CUDA device"):
        cloud.evolve(1.0, 1e-3)

"""The reference's own shipped solidDict files (examples/*/solidDict), read UNCHANGED by the host façade: the drop-in claim for
the configuration surface (reference src/solidcloud.cpp:14-206).  Restart mode (start_time > 0) so no device call is made.
Runs only where the reference tree exists (this container); nothing of it is copied into the repository."""
import math
import os

import numpy as np
import pytest

from sdfibm_b200 import hostapi
from sdfibm_b200.mesh import Mesh

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "examples")), reason="reference tree not present")


def _cloud(rel, tmp_path):
    hostapi.load().sdfibm_host_reset_subiterations()
    mesh = Mesh.hex_block((8, 8, 1), (-4.0, 0.0, -0.5), (1.0, 1.0, 1.0))
    return hostapi.HostCloud(os.path.join(REF, rel), str(tmp_path), mesh, 1.0, start_time=1.0)


def test_flow_past_cylinder_dict(tmp_path):
    c = _cloud("examples/flow_past_cylinder/re200/solidDict", tmp_path)          # C1: one fixed unit circle at the origin
    assert c.n_solids == 1 and c.flags() == (True, True)
    s = c.solids()
    assert np.array_equal(s["pos"][0], (0.0, 0.0, 0.0)) and np.array_equal(s["quat"][0], (1.0, 0.0, 0.0, 0.0))
    assert c.masses()[0] == pytest.approx(math.pi, rel=1e-15)
    c.evolve(1.0, 1e-2)                                                            # frozen motion: nothing moves
    assert np.array_equal(c.solids()["pos"][0], (0.0, 0.0, 0.0)) and np.array_equal(c.solids()["vel"][0], (0.0, 0.0, 0.0))


def test_sedimentation_dict(tmp_path):
    c = _cloud("examples/sedimentation/solidDict", tmp_path)                       # C2: 100 circles r = 0.15, rho = 3, 10 x 10 lattice
    assert c.n_solids == 100 and c.flags() == (True, True)
    s = c.solids()
    xs, ys = np.meshgrid(-1.8 + 0.4 * np.arange(10), 4.2 + 0.4 * np.arange(10))
    want = sorted(zip(np.round(xs.ravel(), 9), np.round(ys.ravel(), 9)))
    got = sorted(zip(np.round(s["pos"][:, 0], 9), np.round(s["pos"][:, 1], 9)))
    assert got == want and np.all(s["pos"][:, 2] == 0.0)
    assert np.allclose(c.masses(), math.pi * 0.15 ** 2 * 3.0, rtol=1e-14)
    # one DEM step under gravity (0, -10, 0) with buoyancy (rho_f = 1): dv = (1 - 1/3) * (-10) * dt per step, 20 sub-iterations
    c.evolve(1.0, 5e-4)
    v = c.solids()["vel"]
    assert np.allclose(v[:, 1], -(2.0 / 3.0) * 10.0 * 5e-4, rtol=1e-12) and np.all(v[:, 0] == 0.0)


def test_falling_ellipse_dict(tmp_path):
    c = _cloud("examples/falling_ellipse/solidDict", tmp_path)                     # C3: Ellipse 0.3 / 0.15 at (0.5, 3.5), z = -45 deg, rho = 2
    assert c.n_solids == 1
    s = c.solids()
    assert np.array_equal(s["pos"][0], (0.5, 3.5, 0.0))
    h = math.radians(-45.0) / 2.0
    assert np.allclose(s["quat"][0], (math.cos(h), 0.0, 0.0, math.sin(h)), atol=1e-15)
    assert c.masses()[0] == pytest.approx(math.pi * 0.3 * 0.15 * 2.0, rel=1e-14)


def test_taylor_couette_dict(tmp_path):
    c = _cloud("examples/taylor_couette/solidDict", tmp_path)                      # C3: Circle r = 0.3, rho = 2, free to spin about z only
    assert c.n_solids == 1 and c.flags() == (True, True)
    assert c.masses()[0] == pytest.approx(math.pi * 0.09 * 2.0, rel=1e-14)
    c.evolve(1.0, 1e-3)                                                            # no gravity, no fluid torque yet: at rest
    s = c.solids()
    assert np.array_equal(s["pos"][0], (0.0, 0.0, 0.0)) and np.array_equal(s["omega"][0], (0.0, 0.0, 0.0))


def test_g1_case_data_equal_the_reference_tool_vof_dict():
    """cases.g1_solids() — the 14 solids behind the hard golden G1 (tests/test_oracle_golden.py, test_gpu_parity.py) — were typed in
    from tool_vof/example/solidDict; here they are checked against that file itself: shape records through the façade's own shape
    plugins, positions and Euler angles through a minimal reading of the `solids` block."""
    import re

    from sdfibm_b200 import cases
    from sdfibm_b200.shapes import quat_from_euler_xyz_deg

    path = os.path.join(REF, "tool_vof/example/solidDict")
    text = re.sub(r"//.*", "", open(path).read())
    shapes, S = cases.g1_solids()
    names = ["circle1", "circle_tail1", "ellipse1", "rectangle1", "plane1"]          # the order of g1_solids()' shape table
    for i, name in enumerate(names):
        rec, _ = hostapi.shape_record(path, name)
        for f in ("tag", "finite", "radiusB"):
            assert rec[f] == shapes[i][f], (name, f)
        assert np.array_equal(rec["p"], shapes[i]["p"]) and np.array_equal(rec["com"], shapes[i]["com"]), name
    block = text[text.index("solids"):]
    entries = re.findall(r"shp_name\s+(\w+)\s*;\s*pos\s*\(([^)]*)\)\s*;\s*euler\s*\(([^)]*)\)\s*;", block)
    assert len(entries) == len(S) == 14
    for i, (shp, pos, euler) in enumerate(entries):
        assert names.index(shp) == S[i]["shape"], i
        assert np.array_equal(np.array(pos.split(), dtype=float), S[i]["pos"]), i
        assert np.allclose(quat_from_euler_xyz_deg(tuple(float(x) for x in euler.split())), S[i]["quat"], rtol=0, atol=0), i

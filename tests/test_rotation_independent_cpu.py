"""3-D rotation arithmetic held to an INDEPENDENT implementation (scipy.spatial.transform.Rotation), not to the repository's own
Foam-free quaternion (host/foamlite.h): the reference builds its orientation with Foam::quaternion(XYZ, euler) =
q_x(ax) q_y(ay) q_z(az) (src/solidcloud.cpp:117 region; OpenFOAM quaternionI.H) — the intrinsic X-Y'-Z'' sequence — and takes
world2local(p) = conjugate(q).transform(p - t) (src/libshape/ishape.h:43-46).  Checked with rotations about ALL THREE axes:
  * the quaternion the host façade parses from `euler (..)` in a solidDict,
  * the python helper the synthetic cases use (shapes.quat_from_euler_xyz_deg),
  * the inside predicate / signed distance of rotated Ellipsoids and Boxes as the oracle (and, via device_math.cuh compiled for the
    host, the product's SDF switch) evaluate them."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

import host_cases as hc
from oracle.oracle_py import eval_points
from sdfibm_b200 import hostapi
from sdfibm_b200.mesh import Mesh
from sdfibm_b200.shapes import make_shape, make_solids, quat_from_euler_xyz_deg

EULERS = [(30.0, -50.0, 70.0), (-120.0, 15.0, 200.0), (5.0, 85.0, -5.0), (90.0, 0.0, 45.0), (0.0, 0.0, 33.0), (179.0, -89.0, 1.0)]


def _scipy_wxyz(e):
    x, y, z, w = Rotation.from_euler("XYZ", e, degrees=True).as_quat()
    return np.array([w, x, y, z])


def _same_rotation(q, p, tol=1e-15):
    return min(np.abs(q - p).max(), np.abs(q + p).max()) <= tol


@pytest.mark.parametrize("euler", EULERS)
def test_euler_xyz_quaternion_matches_scipy(euler, tmp_path):
    assert _same_rotation(np.array(quat_from_euler_xyz_deg(euler)), _scipy_wxyz(euler), 2e-16)
    # the host façade's own parse of `euler` (foamlite quaternion, or OpenFOAM's in the drop-in)
    mesh = Mesh.hex_block((4, 4, 4), (0, 0, 0), (1.0, 1.0, 1.0))
    meta = dict(on_fluid=0, on_twod=0, gravity=(0, 0, 0))
    solids = [dict(shp_name="elo", mot_name="free", mat_name="heavy", pos=(2.0, 2.0, 2.0), euler=euler)]
    path = hc.write_case(tmp_path, meta, solids)
    cloud = hostapi.HostCloud(path, str(tmp_path), mesh, 1.0, start_time=1.0)   # start_time > 0: no initial interact, no device needed
    q = cloud.solids()[0]["quat"]
    cloud.close()
    assert _same_rotation(np.array(q), _scipy_wxyz(euler), 4e-16)


@pytest.mark.parametrize("euler", EULERS)
@pytest.mark.parametrize("kind", ["Ellipsoid", "Box"])
def test_rotated_sdf_matches_an_independent_rotation_matrix(euler, kind):
    rng = np.random.RandomState(int(abs(euler[0]) * 7 + abs(euler[2])))
    abc = np.array([0.5, 0.3, 0.2])
    shapes = np.array([make_shape(kind, radiusa=abc[0], radiusb=abc[1], radiusc=abc[2])])
    S = make_solids(1)
    t = np.array([0.3, -0.2, 0.1])
    S[0]["pos"] = t
    S[0]["quat"] = quat_from_euler_xyz_deg(euler)
    pts = t + rng.uniform(-0.7, 0.7, size=(20000, 3))
    inside, phi = eval_points(shapes, S[0], pts)
    R = Rotation.from_euler("XYZ", euler, degrees=True).as_matrix()      # body -> world
    loc = (pts - t) @ R                                                  # world2local = R^T (p - t)
    if kind == "Ellipsoid":
        g = ((loc / abc) ** 2).sum(axis=1) - 1.0
    else:
        g = (np.abs(loc) - abc).max(axis=1)
    clear = np.abs(g) > 1e-9                                             # points within rounding of the surface decide nothing
    assert clear.sum() > 19000 and (inside[clear] == (g[clear] < 0)).all()
    if kind == "Box":                                                    # exact signed distance (src/libshape/sdf/sdf.h box)
        d = np.abs(loc) - abc
        sd = np.linalg.norm(np.maximum(d, 0.0), axis=1) + np.minimum(d.max(axis=1), 0.0)
        far = np.abs(sd) > 1e-6                                          # the |phi| < 1e-8 filter replaces values at the surface
        assert np.abs(phi[far] - sd[far]).max() <= 1e-13

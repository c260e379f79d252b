"""Randomised oracle-vs-reference comparison on the cases of tests/fuzz_cases.py.  The oracle (both its faithful and its fast list
builder) against the reference's own compiled classes (oracle/_ref): lists and Ct identical, As / Ts / Fs to 1e-13, force / torque
to 1e-11.  8000 box-cell and 1200 general-cell cases were run when this was written (no discrepancy); the test keeps 400 + 160."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fuzz_cases  # noqa: E402
from sdfibm_b200.mesh import Mesh  # noqa: E402
from sdfibm_b200.shapes import make_shape, make_solids, quat_from_euler_xyz_deg  # noqa: E402
from oracle import ref_py  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present")


@pytest.fixture(scope="module", autouse=True)
def _built():
    from sdfibm_b200 import build

    assert build.build_reference_oracle() is not None and ref_py.available()


def _compare(case):
    mesh, two_d, shapes, S, U = case["mesh"], case["two_d"], case["shapes"], case["solids"], case["U"]
    o = Oracle(mesh, two_d)
    mine = o.interact(shapes, S, U, case["dt"], case["rhof"], faithful=True)
    fast = o.interact(shapes, S, U, case["dt"], case["rhof"], faithful=False)
    seeds = np.array([o.nearest_cell(S[i]["pos"]) for i in range(len(S))], dtype=np.int32)
    ref = ref_py.Reference(mesh).interact([ref_py.dict_text_from_record(r) for r in shapes], S, seeds, U, case["dt"], case["rhof"], two_d)
    bad = []
    if not (np.array_equal(ref["list_off"], mine["list_off"]) and np.array_equal(ref["list_cells"], mine["list_cells"])):
        bad.append("lists")
    if not np.array_equal(ref["Ct"], mine["Ct"]):
        bad.append("Ct")
    if not mesh_is_mixed(case) and not (np.array_equal(fast["list_off"], mine["list_off"]) and np.array_equal(fast["list_cells"], mine["list_cells"])):
        bad.append("lists of the fast builder")
    for kf in ("As", "Ts", "Fs"):
        a, b = ref[kf], mine[kf]
        okm = np.isfinite(a)
        if not np.array_equal(okm, np.isfinite(b)):
            bad.append(kf + " finiteness")
        elif okm.any() and np.abs(a[okm] - b[okm]).max() > 1e-13 * max(1.0, np.abs(b[okm]).max()):
            bad.append((kf, float(np.abs(a[okm] - b[okm]).max())))
    okm = np.isfinite(ref["FT"]) & np.isfinite(mine["FT"])
    if okm.any() and np.abs(ref["FT"][okm] - mine["FT"][okm]).max() > 1e-11 * max(1.0, np.abs(mine["FT"][okm]).max()):
        bad.append("FT")
    return bad, int(ref["pairs"])


def mesh_is_mixed(case):
    return case["name"].startswith("mixed3d")      # there the fast (order-free) builder may differ by design


@pytest.mark.parametrize("two_d", [False, True])
def test_random_cases(two_d):
    pairs = 0
    for seed in range(200):
        case = fuzz_cases.box_case(seed, two_d)
        bad, p = _compare(case)
        assert not bad, (case["name"], bad, case["specs"])
        pairs += p
    assert pairs > 5000


@pytest.mark.parametrize("kind", ["skew2d", "skew3d", "prism2d", "mixed3d"])
def test_random_cases_on_general_cells(kind):
    pairs = 0
    for seed in range(40):
        case = fuzz_cases.general_case(seed, kind)
        bad, p = _compare(case)
        assert not bad, (case["name"], bad, case["specs"])
        pairs += p
    assert pairs > 1000


def test_random_collision_steps():
    """UGrid pair order and contact forces: round solids, planes and shapes without a contact function, random and lattice
    positions (exact ties in the cell hash, coincident centres -> NaN normals in both), three grid spacings per case including
    HEAD's delta = -2.  9000 comparisons were run when this was written; 600 are kept."""
    nbad, npairs = 0, 0
    for seed in range(200):
        rng = np.random.RandomState(seed)
        two_d = bool(rng.randint(0,2))
        n = (int(rng.randint(4,12)), int(rng.randint(4,12)), 1 if two_d else int(rng.randint(4,12)))
        h = float(rng.choice([0.5, 1.0, 0.3]))
        mesh = Mesh.hex_block(n, x0=(float(rng.choice([0.0,-1.0])), 0.0, -0.5 if two_d else 0.0), dx=(h,h,1.0 if two_d else h))
        k = int(rng.randint(2, 40))
        r = float(rng.uniform(0.2, 0.8))*h
        round_t = "Circle" if two_d else "Sphere"
        specs=[]
        for i in range(k):
            u = rng.rand()
            if u < 0.75: specs.append((round_t, dict(radius=float(r*rng.choice([1.0,1.0,0.7])))))
            elif u < 0.9: specs.append(("Plane", dict()))
            else: specs.append(("Rectangle", dict(radiusa=r, radiusb=r)) if two_d else ("Ellipsoid", dict(radiusa=r, radiusb=r, radiusc=r*0.8)))
        shapes = np.array([make_shape(t, **kw) for t, kw in specs])
        S = make_solids(k)
        lo, hi = mesh.bounds_min, mesh.bounds_max
        P = rng.uniform(lo - 0.05*(hi-lo), hi + 0.05*(hi-lo), size=(k,3))
        if rng.rand() < 0.3: P = lo + np.round((P-lo)/ (0.5*h)) * 0.5*h   # lattice positions: exact ties in the hash
        if two_d: P[:,2] = 0.0
        S["pos"] = P
        for i in range(k):
            S[i]["quat"] = quat_from_euler_xyz_deg((0,0,float(rng.choice([0,90,180,-90,rng.uniform(-180,180)]))) if two_d else tuple(float(x) for x in rng.choice([0,90,rng.uniform(-180,180)], size=3)))
        S["shape"] = np.arange(k)
        texts = [ref_py.shape_dict_text(t, **kw) for t, kw in specs]
        o = Oracle(mesh, two_d)
        for delta in (float(rng.uniform(0.5, 3.0))*r*2, -2.0, 2.0*r):
            pairs, ft = o.collide(shapes, S, delta)
            rp, rft = ref_py.ref_collide(mesh.bounds_min, mesh.bounds_max, delta, texts, S["pos"], S["quat"])
            npairs += len(rp)
            if not np.array_equal(pairs, rp) or not np.array_equal(ft, rft, equal_nan=True):
                nbad += 1
    assert nbad == 0 and npairs > 5000

"""Randomised oracle-vs-reference comparison on the cases of tests/fuzz_cases.py.  The oracle (both its faithful and its fast list
builder) against the reference's own compiled classes (oracle/_ref): lists and Ct identical, As / Ts / Fs to 1e-13, force / torque
to 1e-11.  68 000 box-cell and 7 200 general-cell cases were run when this was written (no discrepancy); the test keeps 400 + 160."""
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
    HEAD's delta = -2.  69 000 comparisons were run when this was written; 600 are kept."""
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


def _rnd_evolve_case(seed):
    rng = np.random.RandomState(seed)
    f = lambda lo, hi: float(rng.uniform(lo, hi))
    v3 = lambda s=1.0: tuple(float(x) for x in s * rng.standard_normal(3))
    unit = lambda: (lambda v: tuple(float(x) for x in v / np.linalg.norm(v)))(rng.standard_normal(3))
    motions = {
        "mask": dict(type="Motion01Mask", mask="b" + "".join(rng.choice(["0", "1"], size=6))),
        "spin": dict(type="Motion000002", period=f(0.5, 5)),
        "spinfree": dict(type="Motion110002", period=f(0.5, 5)),
        "const": dict(type="Motion222000", u=f(-1, 1), v=f(-1, 1), w=f(-1, 1)),
        "sine": dict(type="MotionSineDirectional", amplitude=f(0.1, 1), period=f(0.5, 3), direction=unit()),
        "rotor": dict(type="MotionRotor", period=f(1, 5), radius=f(0.2, 1), theta0=f(0, 3), selfom=f(-2, 2)),
        "gate": dict(type="MotionOpenClose"),
    }
    forces = {
        "push": dict(type="Constant", force=v3(0.3), torque=v3(0.05)),
        "spring": dict(type="Spring", pivot=v3(1.0), k=f(1, 10), l=f(0.1, 1.0)),
        "mag": dict(type="Magnetic", direction=unit(), A=f(0.1, 1), w=f(0.5, 4)),
    }
    materials = {"heavy": dict(type="General", rho=f(2, 5)), "light": dict(type="General", rho=f(0.5, 1.5))}
    n = int(rng.randint(2, 9))
    solids = []
    for i in range(n):
        s = dict(shp_name=str(rng.choice(["sph", "elo", "box"])), mot_name=str(rng.choice(["free"] + list(motions))),
                 mat_name=str(rng.choice(list(materials))), pos=v3(2.0), vel=v3(0.3), euler=tuple(float(x) for x in rng.uniform(-180, 180, 3)), omega=v3(0.5))
        if rng.rand() < 0.5: s["for_name"] = str(rng.choice(list(forces)))
        solids.append(s)
    g = v3(5.0)
    t0 = f(0.1, 6.0); dt = f(1e-3, 2e-2); steps = int(rng.randint(2, 7))
    return solids, motions, forces, materials, g, t0, dt, steps


def test_random_evolve_cases(tmp_path):
    """SolidCloud::evolve with random plugin parameters (all seven motions, all three forcers, random masks / periods / pivots),
    random states, gravity, step sizes and start times (MotionOpenClose's windows), a fluid force that changes every step:
    oracle/host_oracle.py within 1e-15 of the reference's compiled Solid / libmotion / libforcer, and the shipped C++ façade
    (through a solidDict it parses itself) BIT-IDENTICAL to it.  1 800 cases were run when this was written; 60 are kept."""
    import shutil
    import tempfile

    import host_cases as hc
    from oracle import host_oracle as ho
    from sdfibm_b200 import hostapi

    hostapi.load()
    worst_o = worst_f = 0.0
    for seed in range(60):
        solids, motions, forces, materials, g, t0, dt, steps = _rnd_evolve_case(seed)
        n = len(solids)
        ref = hc.oracle_solids(solids, motions, forces, materials)
        def sdict(name):
            d = dict(hc.SHAPES[name]); return ref_py.shape_dict_text(d.pop("type"), com=d.pop("com", (0.0, 0.0, 0.0)), **d)
        x, q, v, om = hc.state_arrays(ref)
        times = t0 + dt * np.arange(1, steps + 1)
        rng = np.random.RandomState(seed + 10000)
        fluid = 0.05 * rng.standard_normal((steps, n, 6))
        args = ([sdict(s["shp_name"]) for s in solids], [None if s["mot_name"] == "free" else motions[s["mot_name"]] for s in solids],
                [forces[s["for_name"]] if "for_name" in s else None for s in solids], [materials[s["mat_name"]]["rho"] for s in solids], x, q, v, om, times, dt, 20, g)
        out = ref_py.ref_evolve(*args, 1.1, fluid_ft=fluid)
        for step in range(steps):
            for i, s in enumerate(ref): s.ff, s.ft = tuple(fluid[step, i, :3]), tuple(fluid[step, i, 3:])
            ho.evolve(ref, float(times[step]), dt, 20, g, 1.1)
            mine = np.concatenate(hc.state_arrays(ref), axis=1)
            worst_o = max(worst_o, np.abs(mine - out["traj"][step]).max() / max(1.0, np.abs(mine).max()))
        # the façade (DEM mode: rho_f = 0, no fluid force)
        d = tempfile.mkdtemp()
        try:
            path = hc.write_case(d, dict(on_fluid=0, on_twod=0, gravity=g), solids, motions=motions, forces=forces, materials=materials)
            hostapi.load().sdfibm_host_reset_subiterations()
            cloud = hostapi.HostCloud(path, d, Mesh.hex_block((2, 2, 2)), rho_fluid=1.0, start_time=t0)
            st = cloud.solids()
            out2 = ref_py.ref_evolve(args[0], args[1], args[2], args[3], st["pos"], st["quat"], st["vel"], st["omega"], times, dt, 20, g, 0.0)
            for step in range(steps):
                cloud.evolve(float(times[step]), dt)
                got = cloud.solids()
                mine = np.concatenate([got["pos"], got["quat"], got["vel"], got["omega"]], axis=1)
                worst_f = max(worst_f, np.abs(mine - out2["traj"][step]).max() / max(1.0, np.abs(mine).max()))
            cloud.close()
        finally:
            shutil.rmtree(d)
    assert worst_o <= 1e-15 and worst_f == 0.0, (worst_o, worst_f)


def test_random_shapes_mass_properties_and_point_evaluation(tmp_path):
    """The nine shapes with random parameters and com offsets: volume, volumeINV, radiusB, moments (the façade's shape plugins, read
    from a solidDict) and phi01 / phi at random world points for a random pose (the ONE device / host evaluation switch,
    csrc/device_math.cuh, through the façade and through the oracle) against the reference's own shape constructors and
    Solid::phi01 / Solid::phi — bit for bit."""
    import host_cases as hc
    from oracle.oracle_py import eval_points
    from sdfibm_b200 import hostapi

    rng = np.random.RandomState(4)
    f = lambda lo, hi: float(rng.uniform(lo, hi))
    com = lambda two_d: (f(-0.1, 0.1), f(-0.1, 0.1), 0.0 if two_d else f(-0.1, 0.1))
    n_checked = 0
    for rep in range(12):
        shapes = {
            "circ": dict(type="Circle", radius=f(0.1, 1), com=com(True)),
            "sph": dict(type="Sphere", radius=f(0.1, 1), com=com(False)),
            "ell": dict(type="Ellipse", radiusa=f(0.1, 1), radiusb=f(0.1, 1), com=com(True)),
            "elo": dict(type="Ellipsoid", radiusa=f(0.1, 1), radiusb=f(0.1, 1), radiusc=f(0.1, 1)),
            "rect": dict(type="Rectangle", radiusa=f(0.1, 1), radiusb=f(0.1, 1), com=com(True)),
            "box": dict(type="Box", radiusa=f(0.1, 1), radiusb=f(0.1, 1), radiusc=f(0.1, 1), com=com(False)),
            "tail": dict(type="Circle_Tail", radius=f(0.1, 0.6), ratio=f(0.5, 3), thickness=f(0.02, 0.2), com=com(True)),
            "twotail": dict(type="Circle_TwoTail", radius=f(0.1, 0.6), ratio=f(0.5, 3), thickness=f(0.02, 0.2), com=com(True)),
            "plane": dict(type="Plane"),
        }
        d = tmp_path / f"r{rep}"
        d.mkdir()
        path = hc.write_case(d, dict(on_fluid=0, on_twod=0, gravity=(0.0, 0.0, 0.0)), [], shapes=shapes)
        pos = rng.uniform(-0.3, 0.3, 3)
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        pts = rng.uniform(-1.5, 1.5, size=(500, 3))
        pts[:50] = pos                                        # the centre itself (ellipse / ellipsoid: -1/0)
        for name, spec in shapes.items():
            k = dict(spec)
            text = ref_py.shape_dict_text(k.pop("type"), com=k.pop("com", (0.0, 0.0, 0.0)), **k)
            ref = ref_py.ref_shape_props(text, pos, q, pts)
            rec, props = hostapi.shape_record(path, name)
            assert props["volume"] == ref["volume"] and props["volumeINV"] == ref["volumeINV"] and props["radiusB"] == ref["radiusB"], name
            assert bool(rec["finite"]) == ref["finite"] and np.array_equal(rec["com"], ref["com"]), name
            assert np.array_equal(np.asarray(props["moi"]), ref["moi"]), name
            inside, phi = hostapi.shape_eval(path, name, pos, q, pts)
            assert np.array_equal(inside, ref["inside"]) and np.array_equal(phi, ref["phi"], equal_nan=True), name
            from sdfibm_b200.shapes import make_solids

            S = make_solids(1)
            S[0]["pos"], S[0]["quat"] = pos, q
            oi, op = eval_points(np.array([rec]), S[0], pts)
            assert np.array_equal(oi, ref["inside"]) and np.array_equal(op, ref["phi"], equal_nan=True), name
            n_checked += 1
    assert n_checked == 108

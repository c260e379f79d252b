"""Host façade on the GPU: the whole per-step sequence main.cpp runs on SolidCloud (interact -> U -= Fs dt -> evolve ->
saveState -> fixInternal, reference src/main.cpp:66-88) through the C++ SolidCloud, against the oracle chain
(oracle interact / collide / fixInternal + the Python restatement of evolve) on the same solidDict."""
import os
import subprocess

import numpy as np
import pytest

import host_cases as hc
from oracle import host_oracle as ho
from oracle.oracle_py import Oracle
from sdfibm_b200 import cases, hostapi
from sdfibm_b200.mesh import Mesh

pytestmark = pytest.mark.gpu


def _sedimentation(n_side=5):
    """Scaled-down examples/sedimentation: circles r = 0.15 on a lattice in a box closed by four plane walls."""
    solids = []
    for j in range(n_side):
        for i in range(n_side):
            solids.append(dict(shp_name="circ", mot_name="free", mat_name="heavy", pos=(-0.8 + 0.4 * i + 0.03 * ((i + j) % 3), 2.2 + 0.36 * j, 0.0),
                               vel=(0.05 * (i - 2), -0.1 * j, 0.0), omega=(0.0, 0.0, 0.3 * (i - j))))
    walls = [((0.0, 0.0, 0.0), 0.0), ((0.0, 4.0, 0.0), 180.0), ((-2.0, 2.0, 0.0), -90.0), ((2.0, 2.0, 0.0), 90.0)]
    for p, ez in walls:
        solids.append(dict(shp_name="plane", mot_name="fixed", mat_name="heavy", pos=p, euler=(0.0, 0.0, ez)))
    return solids


def test_coupled_steps_match_oracle_chain(tmp_path):
    solids = _sedimentation()
    motions = dict(hc.MOTIONS, fixed=dict(type="Motion01Mask", mask="b000000"))
    meta = dict(on_fluid=1, on_twod=1, gravity=(0.0, -10.0, 0.0), collision_delta=0.33)
    path = hc.write_case(tmp_path, meta, solids, motions=motions)
    mesh = Mesh.hex_block((200, 200, 1), (-2.0, 0.0, -0.5), (0.02, 0.02, 1.0))
    U0 = cases.taylor_green(mesh.cc, 4.0)
    rhof, dt = 1.0, 5e-4
    hostapi.load().sdfibm_host_reset_subiterations()
    cloud = hostapi.HostCloud(path, str(tmp_path), mesh, rhof, start_time=0.0, U_init=U0)
    assert os.path.exists(os.path.join(str(tmp_path), "0_As"))            # initialCorrect wrote As (src/solidcloud.cpp:276-281)

    table, index = hc.shape_table(solids)
    ref = hc.oracle_solids(solids, motions=motions)
    o = Oracle(mesh, True)
    U = U0.copy()
    # initialCorrect: interact(0, dt = 1) — only As survives, the fluid force is overwritten before the first evolve
    r0 = o.interact(table, ho.records(ref, index), U, 1.0, rhof)
    assert np.abs(cloud.field("As") - r0["As"]).max() <= 1e-12

    def collide(ss):
        pairs, ft = o.collide(table, ho.records(ss, index), meta["collision_delta"])
        return ft if len(pairs) else None

    t = 0.0
    n_pairs_seen = 0
    for step in range(3):
        t += dt
        cloud.interact(t, dt)
        r = o.interact(table, ho.records(ref, index), U, dt, rhof)
        for k in ("As", "Ts", "Ct"):
            assert np.abs(cloud.field(k) - r[k]).max() <= 1e-9 * max(1.0, np.abs(r[k]).max()), (step, k)
        assert np.abs(cloud.field("Fs") - r["Fs"]).max() <= 1e-9 * np.abs(r["Fs"]).max(), step
        _, fluid = cloud.forces()
        assert np.abs(fluid - r["FT"]).max() <= 1e-9 * np.abs(r["FT"]).max(), step
        for i, s in enumerate(ref):
            s.ff, s.ft = tuple(r["FT"][i, :3]), tuple(r["FT"][i, 3:])
        # main.cpp:70
        cloud.field("U")[:] = cloud.field("U") - cloud.field("Fs") * dt
        U = U - r["Fs"] * dt
        cloud.evolve(t, dt)
        n_pairs_seen += len(o.collide(table, ho.records(ref, index), meta["collision_delta"])[0])
        ho.evolve(ref, t, dt, 20, meta["gravity"], rhof, collide=collide)
        cloud.save_state()
        cloud.fix_internal(dt)
        U = o.fix_internal(table, ho.records(ref, index), r["Ct"], U)
        x, q, v, om = hc.state_arrays(ref)
        got = cloud.solids()
        for a, b in ((got["pos"], x), (got["quat"], q), (got["vel"], v), (got["omega"], om)):
            assert np.abs(a - b).max() <= 1e-8 * max(1.0, np.abs(b).max()), step
        assert np.abs(cloud.field("U") - U).max() <= 1e-8 * np.abs(U).max(), step
    assert n_pairs_seen > 0                                                  # the collision step was exercised
    # planes are held by their motion mask
    assert np.array_equal(got["pos"][-4:], [s["pos"] for s in solids[-4:]])


def test_taylor_couette_example_through_the_facade(tmp_path):
    """BASELINE config 3 as the reference sets it up — examples/taylor_couette: the five-block O-grid of its blockMeshDict
    (sdfibm_b200.meshgen), the solidDict's one Circle r = 0.3 at the origin, material rho 2, Motion01Mask b000001 (free to spin about
    z only: the torque-coupled motion), no gravity — three coupled steps through the C++ SolidCloud against the oracle chain."""
    from sdfibm_b200.meshgen import ogrid_taylor_couette

    motions = dict(onlyzrot1=dict(type="Motion01Mask", mask="b000001"))
    materials = dict(mat1=dict(type="General", rho=2.0))
    solids = [dict(shp_name="circle_tc", mot_name="onlyzrot1", mat_name="mat1", pos=(0.0, 0.0, 0.0), vel=(0.0, 0.0, 0.0), euler=(0.0, 0.0, 0.0))]
    meta = dict(on_fluid=1, on_twod=1, gravity=(0.0, 0.0, 0.0))
    path = hc.write_case(tmp_path, meta, solids, motions=motions, materials=materials)
    mesh = ogrid_taylor_couette(30)
    # a swirling initial field (the outer cylinder drives the flow in the example): it exerts a torque on the inner circle
    r2 = mesh.cc[:, 0] ** 2 + mesh.cc[:, 1] ** 2
    U0 = np.stack([-mesh.cc[:, 1] * (0.5 + r2), mesh.cc[:, 0] * (0.5 + r2), np.zeros(mesh.n_cells)], axis=1)
    rhof, dt = 1.0, 1e-3
    hostapi.load().sdfibm_host_reset_subiterations()
    cloud = hostapi.HostCloud(path, str(tmp_path), mesh, rhof, start_time=0.0, U_init=U0)
    table, index = hc.shape_table(solids)
    ref = hc.oracle_solids(solids, motions=motions, materials=materials)
    o = Oracle(mesh, True)
    U = U0.copy()
    t = 0.0
    for step in range(3):
        t += dt
        cloud.interact(t, dt)
        r = o.interact(table, ho.records(ref, index), U, dt, rhof)
        assert np.array_equal(cloud.field("Ct"), r["Ct"]), step
        for k in ("As", "Ts", "Fs"):
            assert np.abs(cloud.field(k) - r[k]).max() <= 1e-9 * max(1.0, np.abs(r[k]).max()), (step, k)
        _, fluid = cloud.forces()
        assert np.abs(fluid - r["FT"]).max() <= 1e-9 * np.abs(r["FT"]).max(), step
        assert abs(r["FT"][0, 5]) > 1e-3                                        # the swirl turns the circle
        for i, sd in enumerate(ref):
            sd.ff, sd.ft = tuple(r["FT"][i, :3]), tuple(r["FT"][i, 3:])
        cloud.field("U")[:] = cloud.field("U") - cloud.field("Fs") * dt       # main.cpp:70
        U = U - r["Fs"] * dt
        cloud.evolve(t, dt)
        ho.evolve(ref, t, dt, 20, meta["gravity"], rhof)
        cloud.save_state()
        cloud.fix_internal(dt)
        U = o.fix_internal(table, ho.records(ref, index), r["Ct"], U)
        x, q, v, om = hc.state_arrays(ref)
        got = cloud.solids()
        for a_, b_ in ((got["pos"], x), (got["quat"], q), (got["vel"], v), (got["omega"], om)):
            assert np.abs(a_ - b_).max() <= 1e-8 * max(1.0, np.abs(b_).max()), step
        assert np.abs(cloud.field("U") - U).max() <= 1e-8 * np.abs(U).max(), step
    got = cloud.solids()
    assert np.array_equal(got["pos"][0], (0.0, 0.0, 0.0)) and got["omega"][0, 2] != 0.0 and not got["omega"][0, :2].any()   # only z rotation is free
    cloud.close()


def test_standalone_runner_3d(tmp_path):
    """sdfibm_b200_run: the main.cpp-shaped loop without OpenFOAM on a small 3-D case."""
    solids = [dict(shp_name="sph", mot_name="free", mat_name="heavy", pos=(1.6, 1.7, 1.5), omega=(0.0, 0.0, 0.5)),
              dict(shp_name="elo", mot_name="free", mat_name="light", pos=(3.1, 2.2, 2.6), euler=(10.0, 20.0, 30.0))]
    hc.write_case(tmp_path, dict(on_fluid=1, on_twod=0, gravity=(0.0, 0.0, -9.81)), solids)
    with open(os.path.join(str(tmp_path), "runDict"), "w") as f:
        f.write("mesh { cells (48 48 48); origin (0 0 0); spacing (0.1 0.1 0.1); }\n"
                "fluid { rho 1.0; U (0.1 0 0); }\ntime { deltaT 1e-3; nSteps 3; startTime 0; }\n")
    hostapi.load()
    r = subprocess.run([hostapi.RUNNER_PATH, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "ran 3 steps, 2 solids" in r.stdout
    rows = open(os.path.join(str(tmp_path), "cloud.out")).read().strip().split("\n")
    assert len(rows) == 2 * 4 and all(len(x.split()) == 19 for x in rows)
    vol = float(r.stdout.split("solid volume")[1])
    exact = 4.0 / 3.0 * np.pi * (0.3 ** 3 + 0.5 * 0.45 * 0.4)
    assert -0.05 < vol / exact - 1.0 < 0.0                                   # the algorithm's known negative bias
    assert os.path.exists(os.path.join(str(tmp_path), "solidDict.restart"))
    log = open(os.path.join(str(tmp_path), "cloud.log")).read()
    assert "FSI took" in log
    # the restart layout of src/main.cpp:25-36,93-100: <endTime>/solidDict is written and, on a restart from that time, read
    assert os.path.exists(os.path.join(str(tmp_path), "0.003", "solidDict"))
    with open(os.path.join(str(tmp_path), "runDict"), "w") as f:
        f.write("mesh { cells (48 48 48); origin (0 0 0); spacing (0.1 0.1 0.1); }\n"
                "fluid { rho 1.0; U (0.1 0 0); }\ntime { deltaT 1e-3; nSteps 2; startTime 0.003; parallel 1; }\n")
    os.makedirs(os.path.join(str(tmp_path), "processor0", "0.003"))
    os.replace(os.path.join(str(tmp_path), "0.003", "solidDict"), os.path.join(str(tmp_path), "processor0", "0.003", "solidDict"))
    os.remove(os.path.join(str(tmp_path), "solidDict"))          # only the time-directory copy is left: it must be the one read
    r2 = subprocess.run([hostapi.RUNNER_PATH, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0, r2.stderr
    assert "ran 2 steps, 2 solids" in r2.stdout and os.path.exists(os.path.join(str(tmp_path), "processor0", "0.005", "solidDict"))


def test_mean_field_output(tmp_path):
    """meta { on_meanfield 1; sampler <shape>; }: saveState appends one row of 3 N means to meanfield.out
    (reference src/solidcloud.cpp:29-34,303-313,587-590) and leaves Ct / fixInternal alone."""
    solids = [dict(shp_name="sph", mot_name="free", mat_name="heavy", pos=(1.6, 1.7, 1.5)),
              dict(shp_name="box", mot_name="free", mat_name="light", pos=(3.1, 2.2, 2.6), euler=(10.0, 20.0, 30.0))]
    meta = dict(on_fluid=1, on_twod=0, gravity=(0.0, 0.0, -9.81), on_meanfield=1, sampler="elo")
    path = hc.write_case(tmp_path, meta, solids)
    mesh = Mesh.hex_block((48, 48, 48), (0, 0, 0), (0.1, 0.1, 0.1))
    U0 = cases.taylor_green(mesh.cc, 4.8)
    hostapi.load().sdfibm_host_reset_subiterations()
    cloud = hostapi.HostCloud(path, str(tmp_path), mesh, 1.0, start_time=0.0, U_init=U0)
    cloud.interact(1e-3, 1e-3)
    ct = cloud.field("Ct").copy()
    cloud.save_state()
    rows = open(os.path.join(str(tmp_path), "meanfield.out")).read().strip().split("\n")
    assert len(rows) == 1 and len(rows[0].split()) == 6
    got = np.array(rows[0].split(), dtype=float).reshape(2, 3)
    assert np.allclose(got, cloud.mean_field(), rtol=1e-6)
    # against the oracle: alpha of the sampler ellipsoid placed at each solid
    table, index = hc.shape_table(solids)
    o = Oracle(mesh, False)
    ref = hc.oracle_solids(solids)
    recs = ho.records(ref, [list(hc.SHAPES).index("elo")] * 2)
    for i in range(2):
        a = o.interact(table, recs[i:i + 1], U0, 1.0, 1.0)["Ts"]
        w = a * mesh.V
        assert np.abs(cloud.mean_field()[i] - w @ U0 / w.sum()).max() <= 1e-10
    assert np.array_equal(cloud.field("Ct"), ct)

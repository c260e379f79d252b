"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): candidate lists and collision pair sets bit-exact; As / Fs (and Ts, Ct)
within 1e-12 relative; per-solid force / torque within 1e-10 relative to the sum of |terms|.
"""
import os

import numpy as np
import pytest

from oracle.oracle_py import Oracle
from sdfibm_b200 import cases
from sdfibm_b200.context import Context

pytestmark = pytest.mark.gpu

REL_FIELD = 1e-12
REL_FORCE = 1e-10


def _term_scale(case, off, cells):
    """Per-solid upper bound of sum |terms| of the force/torque sums (alpha <= 1)."""
    m, S = case["mesh"], case["solids"]
    n = len(S)
    scale = np.zeros((n, 6))
    for s in range(n):
        cs = cells[off[3 * s]:off[3 * s + 3]]
        if len(cs) == 0:
            continue
        r = m.cc[cs] - S[s]["pos"]
        us = S[s]["vel"] + np.cross(S[s]["omega"], r)
        f = np.abs(case["U"][cs] - us) * m.V[cs, None] / case["dt"] * case["rhof"]
        scale[s, :3] = f.sum(axis=0)
        scale[s, 3:] = (np.linalg.norm(r, axis=1)[:, None] * np.linalg.norm(f, axis=1)[:, None]).sum()
    return scale


def assert_fields_close(got, ref, rel=REL_FIELD):
    """As / Ts / Fs cell by cell (component by component) to `rel` RELATIVE to the reference's own value — sliver fractions
    included; the only absolute allowance is 1e-300 (exact zeros)."""
    for k in ("As", "Ts", "Fs"):
        a, b = got[k], ref[k]
        bad = np.abs(a - b) > rel * np.abs(b) + 1e-300
        if bad.any():
            i = np.argmax(np.abs(a - b) / (np.abs(b) + 1e-300) * bad)
            raise AssertionError((k, int(bad.sum()), float(a.flat[i]), float(b.flat[i])))


def run_both(case, cell_slots=None):
    o = Oracle(case["mesh"], case["two_d"])
    ref = o.interact(case["shapes"], case["solids"], case["U"], case["dt"], case["rhof"])
    ctx = Context(0, cell_slots=cell_slots)
    ctx.set_mesh(case["mesh"], case["two_d"])
    ctx.set_shapes(case["shapes"])
    got = ctx.interact(case["solids"], case["U"], case["dt"], case["rhof"])
    return o, ref, ctx, got


def check_parity(case, o, ref, ctx, got):
    off, cells = ctx.candidate_lists()
    assert np.array_equal(off, ref["list_off"]), "candidate list sizes differ"
    assert np.array_equal(cells, ref["list_cells"]), "candidate lists differ"
    counts = ctx.candidate_counts()
    n = len(case["solids"])
    assert counts == [int((off[1::3][:n] - off[0::3][:n]).sum()), int((off[2::3][:n] - off[1::3][:n]).sum()),
                      int((off[3::3][:n] - off[2::3][:n]).sum())]
    assert np.array_equal(got["Ct"], ref["Ct"])
    assert_fields_close(got, ref)
    scale = _term_scale(case, ref["list_off"], ref["list_cells"])
    err = np.abs(got["FT"] - ref["FT"])
    assert (err <= REL_FORCE * np.maximum(scale, np.abs(ref["FT"])) + 1e-300).all(), err.max()
    # fixInternal with moved solids (solid state AFTER evolve, Ct of the last interact)
    S2 = case["solids"].copy()
    S2["vel"] += 0.05
    S2["omega"] *= 1.1
    S2["pos"] += 0.01
    U2 = case["U"] * 0.9
    assert np.array_equal(ctx.fix_internal(S2, U2), o.fix_internal(case["shapes"], S2, ref["Ct"], U2))


@pytest.mark.parametrize("name", ["c1", "c2", "c2_walls", "skewed2d", "mixed3d", "prism2d", "c4_small"])
def test_parity_synthetic(name):
    case = {
        "c1": cases.case_c1,
        "c2": cases.case_c2,
        "c2_walls": lambda: cases.case_c2(with_walls=True),
        "skewed2d": cases.case_skewed_2d,
        "mixed3d": cases.case_mixed3d,
        "prism2d": cases.case_prism2d,
        "c4_small": lambda: cases.case_c4(n=48, n_solids=50, n_side=4),
    }[name]()
    o, ref, ctx, got = run_both(case, cell_slots=8 if name == "mixed3d" else None)
    check_parity(case, o, ref, ctx, got)
    assert sum(ctx.candidate_counts()) > 0


def _renumber_cells(mesh, perm):
    """The same polyMesh with cell c relabelled perm[c] (faces whose owner would exceed their neighbour are flipped)."""
    from sdfibm_b200.mesh import Mesh

    owner, neigh = perm[mesh.owner].astype(np.int32), perm[mesh.neighbour].astype(np.int32)
    fp_off, fp = mesh.fp_off.copy(), mesh.fp.copy()
    for f in np.nonzero(owner[: len(neigh)] > neigh)[0]:
        owner[f], neigh[f] = neigh[f], owner[f]
        fp[fp_off[f]:fp_off[f + 1]] = fp[fp_off[f]:fp_off[f + 1]][::-1].copy()
    return Mesh.from_polymesh(mesh.points.copy(), fp_off, fp, owner, neigh)


@pytest.mark.parametrize("name", ["c4_small", "mixed3d"])
def test_parity_scrambled_cell_numbering(name):
    """The library renumbers cells into tile order internally: results must not depend on the caller's numbering being
    structured.  A random relabelling of the cells is the worst case for that mapping (every gather / scatter is scattered)."""
    base = cases.case_c4(n=40, n_solids=27, n_side=3) if name == "c4_small" else cases.case_mixed3d()
    rng = np.random.RandomState(7)
    perm = rng.permutation(base["mesh"].n_cells).astype(np.int32)
    mesh = _renumber_cells(base["mesh"], perm)
    U = np.empty_like(base["U"])
    U[perm] = base["U"]
    assert np.allclose(mesh.cc[perm], base["mesh"].cc, rtol=0, atol=1e-12)
    case = dict(base, mesh=mesh, U=U)
    o, ref, ctx, got = run_both(case, cell_slots=8 if name == "mixed3d" else None)
    check_parity(case, o, ref, ctx, got)
    # and the relabelled result is the original one, cell by cell: the classification exactly; the fractions only roughly, because
    # the reference's apex construction starts from a face's FIRST vertex (geometrictools.cpp:74-96) and flipped faces start elsewhere
    o0 = Oracle(base["mesh"], base["two_d"]).interact(base["shapes"], base["solids"], base["U"], base["dt"], base["rhof"])
    assert np.array_equal(got["Ct"][perm], o0["Ct"])
    assert np.abs(got["As"][perm] - o0["As"]).max() <= 0.1


def test_parity_graded_box_cells():
    """Axis-aligned cells of varying size (a graded blockMesh): k_classify's nearest / farthest-corner bounds use per-cell boxes."""
    from sdfibm_b200.mesh import Mesh

    base = cases.case_c4(n=40, n_solids=27, n_side=3)
    pts = base["mesh"].points.copy()
    for d, (a, w) in enumerate([(1.5, 0.31), (-1.1, 0.23), (0.9, 0.17)]):
        pts[:, d] = pts[:, d] + a * np.sin(w * pts[:, d])          # monotone map: cells stay boxes, sizes vary by ~2x
    mesh = Mesh.hex_block_with_points((40, 40, 40), pts)
    assert (np.diff(np.sort(np.unique(pts[:, 0]))) > 0.3).all()
    case = dict(base, mesh=mesh)
    o, ref, ctx, got = run_both(case)
    check_parity(case, o, ref, ctx, got)
    assert sum(ctx.candidate_counts()) > 1000


def test_small_balls_skip_the_connectivity_certificate():
    """Balls barely larger than a cell diagonal, at random offsets: k_solid_prepare declares their cell sets connected without the
    certificate pass; the oracle's real flood fill must agree cell for cell (a wrong proof would show up as missing cells)."""
    from sdfibm_b200.mesh import Mesh
    from sdfibm_b200.shapes import make_shape, make_solids

    rng = np.random.RandomState(11)
    n = 28
    mesh = Mesh.hex_block((n, n, n))
    shapes = np.array([make_shape("Sphere", radius=1.7321), make_shape("Sphere", radius=1.9), make_shape("Sphere", radius=1.0)])
    S = make_solids(150)
    S["pos"] = rng.uniform(4.0, n - 4.0, size=(150, 3))
    S["pos"][:20] = np.round(S["pos"][:20] * 2) / 2          # centres on vertices / face centres: many exact ties
    S["shape"] = rng.randint(0, 3, size=150).astype(np.int32)   # radius 1.0 is below the bound: certificate path
    S["vel"] = 0.1 * rng.standard_normal((150, 3))
    U = rng.standard_normal((mesh.n_cells, 3))
    case = dict(name="small_balls", mesh=mesh, two_d=False, shapes=shapes, solids=S, U=U, dt=1e-3, rhof=1.0)
    o, ref, ctx, got = run_both(case, cell_slots=24)
    check_parity(case, o, ref, ctx, got)


def test_parity_mixed_hex_prism_polyhedron_mesh():
    """A mesh that is mostly hexahedra with some prisms and 7-faced polyhedra (what snappyHexMesh-like tools produce): the
    hexahedra keep the warp-cooperative kernel, the rest are queued separately for the general-polyhedron kernel."""
    from mixed_mesh import mixed_hex_prism_mesh
    from sdfibm_b200.shapes import make_shape, make_solids, quat_from_euler_xyz_deg

    mesh = mixed_hex_prism_mesh(14)
    assert sorted(set(np.diff(mesh.cf_off))) == [5, 6, 7]
    rng = np.random.RandomState(3)
    shapes = np.array([make_shape("Sphere", radius=3.2), make_shape("Ellipsoid", radiusa=3.5, radiusb=2.5, radiusc=2.0),
                       make_shape("Box", radiusa=2.2, radiusb=1.6, radiusc=2.7)])
    S = make_solids(6)
    S["pos"] = [(5.3, 6.1, 5.7), (9.9, 4.2, 8.7), (4.1, 10.2, 9.6), (10.4, 10.1, 4.3), (7.0, 7.0, 7.0), (12.5, 2.0, 12.0)]
    S["shape"] = [0, 1, 2, 0, 1, 2]
    for i, e in enumerate([(0, 0, 0), (20, 40, -15), (35, -10, 60), (0, 0, 0), (-70, 15, 5), (10, 20, 30)]):
        S[i]["quat"] = quat_from_euler_xyz_deg(e)
    S["vel"] = 0.1 * rng.standard_normal((6, 3))
    S["omega"] = 0.05 * rng.standard_normal((6, 3))
    U = rng.standard_normal((mesh.n_cells, 3))
    case = dict(name="mixed_cells", mesh=mesh, two_d=False, shapes=shapes, solids=S, U=U, dt=1e-3, rhof=1.3)
    # On a mesh that mixes cell types the reference's ALL_INSIDE test compares a cell's inside-vertex count with the vertex count
    # of whichever neighbour discovered it first (cellenumerator.cpp:25), i.e. with the flood fill's visiting order.
    # The library uses the cell's own count; the oracle's order-free variant is the specification here, and the faithful run
    # differs from it only in the TYPE of a few cells next to a cell of another kind.
    o = Oracle(mesh, False)
    ref = o.interact(shapes, S, U, 1e-3, 1.3, own_vertex_count=True)
    from sdfibm_b200.capi import SdfibmError
    with pytest.raises(SdfibmError, match="visiting order"):      # refused unless the order-free rule is accepted explicitly
        Context(0, cell_slots=8).set_mesh(mesh, False)
    ctx = Context(0, cell_slots=8, allow_order_free=True)
    ctx.set_mesh(mesh, False)
    ctx.set_shapes(shapes)
    got = ctx.interact(S, U, 1e-3, 1.3)
    check_parity(case, o, ref, ctx, got)
    assert sum(ctx.candidate_counts()) > 500
    faithful = o.interact(shapes, S, U, 1e-3, 1.3)
    assert faithful["list_off"][-1] == ref["list_off"][-1]                      # same member cells ...
    differ = np.nonzero(faithful["Ct"] != ref["Ct"])[0]
    assert 0 < len(differ) < 0.05 * ref["list_off"][-1]                         # ... a few of them typed differently
    ok = np.isfinite(faithful["As"])       # (the faithful run divides by |grad f| = 0 at an ellipsoid centre that sits on a vertex)
    assert np.abs(faithful["As"][ok] - ref["As"][ok]).max() < 1e-12             # a fully inside cell has fraction 1 either way


def test_graph_replay_follows_changing_arguments():
    """The device-resident entry replays a captured CUDA graph: dt, rhof, the solid states and the buffers may change between steps."""
    import torch

    case = cases.case_c4(n=32, n_solids=8, n_side=2)
    o = Oracle(case["mesh"], case["two_d"])
    ctx = Context(0)
    ctx.set_mesh(case["mesh"], case["two_d"])
    ctx.set_shapes(case["shapes"])
    nC, nS = case["mesh"].n_cells, len(case["solids"])
    dev = torch.device("cuda", 0)
    for step, (dt, rhof, shift, fresh) in enumerate([(1e-3, 1.0, 0.0, False), (2e-3, 1.5, 0.3, False), (5e-4, 0.7, -0.2, True), (5e-4, 0.7, 0.1, False)]):
        S = case["solids"].copy()
        S["pos"] += shift
        U = case["U"] * (1.0 + 0.1 * step)
        if step == 0 or fresh:
            buf = [torch.empty(nC * k, dtype=torch.float64, device=dev) for k in (3, 1, 3, 1, 1)] + [torch.empty(nS * 6, dtype=torch.float64, device=dev)]
        buf[0].copy_(torch.from_numpy(U.reshape(-1)))
        torch.cuda.synchronize()
        ctx.interact_device(S, *[b.data_ptr() for b in buf[:1]], dt, rhof, *[b.data_ptr() for b in buf[1:]])
        ref = o.interact(case["shapes"], S, U, dt, rhof)
        As, Fs, Ct, FT = buf[1].cpu().numpy(), buf[2].cpu().numpy().reshape(-1, 3), buf[4].cpu().numpy(), buf[5].cpu().numpy().reshape(-1, 6)
        assert np.array_equal(Ct, ref["Ct"])
        assert np.abs(As - ref["As"]).max() <= 1e-12
        assert np.abs(Fs - ref["Fs"]).max() <= 1e-12 * max(1.0, np.abs(ref["Fs"]).max())
        assert np.abs(FT - ref["FT"]).max() <= 1e-10 * max(1.0, np.abs(ref["FT"]).max())


def test_device_resident_solids_entry():
    """sdfibm_interact_device_solids: the solid records are already on the device (multi-GPU hosts gather them over NVLink)."""
    import torch
    from sdfibm_b200 import capi, parallel

    case = cases.case_mixed3d()
    o = Oracle(case["mesh"], case["two_d"])
    ref = o.interact(case["shapes"], case["solids"], case["U"], case["dt"], case["rhof"])
    ctx = Context(0, cell_slots=8)
    ctx.set_mesh(case["mesh"], case["two_d"])
    ctx.set_shapes(case["shapes"])
    dev = torch.device("cuda", 0)
    nC, nS = case["mesh"].n_cells, len(case["solids"])
    solids = capi.pinned_like(np.ascontiguousarray(case["solids"], dtype=capi.SOLID_DTYPE))
    rep = parallel.ReplicatedSolids(nS, solids.dtype.itemsize, dev)          # world size 1: a plain upload
    ext = torch.cuda.ExternalStream(ctx.stream_ptr(), device=dev)
    dU = torch.from_numpy(case["U"]).to(dev)
    buf = [torch.empty(nC * k, dtype=torch.float64, device=dev) for k in (1, 3, 1, 1)] + [torch.empty(nS * 6, dtype=torch.float64, device=dev)]
    torch.cuda.synchronize()
    with torch.cuda.stream(ext):
        ctx.interact_device_solids(rep.refresh(solids), nS, dU.data_ptr(), case["dt"], case["rhof"], *[b.data_ptr() for b in buf])
    assert np.array_equal(buf[3].cpu().numpy(), ref["Ct"])
    assert np.abs(buf[0].cpu().numpy() - ref["As"]).max() <= 1e-12
    assert np.abs(buf[4].cpu().numpy().reshape(-1, 6) - ref["FT"]).max() <= 1e-10 * max(1.0, np.abs(ref["FT"]).max())
    off, cells = ctx.candidate_lists()
    assert np.array_equal(off, ref["list_off"]) and np.array_equal(cells, ref["list_cells"])


def test_unknown_shape_index_is_an_error():
    case = cases.case_c4(n=32, n_solids=4, n_side=2)
    ctx = Context(0)
    ctx.set_mesh(case["mesh"], case["two_d"])
    ctx.set_shapes(case["shapes"])
    S = case["solids"].copy()
    S["shape"][2] = 7
    with pytest.raises(Exception, match="unknown shape index"):
        ctx.interact(S, case["U"], case["dt"], case["rhof"])


def test_parity_g1_and_golden(m1_points, g1_alpha):
    case = cases.case_g1(m1_points)
    o, ref, ctx, got = run_both(case)
    check_parity(case, o, ref, ctx, got)
    assert np.abs(got["As"] - g1_alpha).max() <= 1e-15      # the GPU itself against the reference's golden


def test_parity_c3_shipped_mesh(m2_points):
    case = cases.case_c3(m2_points)
    check_parity(case, *run_both(case))


def test_parity_disconnected_component_replay():
    """a vertex-inside set that is not face connected must be cut to the seed's component."""
    case = cases.case_disconnected()
    o, ref, ctx, got = run_both(case)
    assert ctx.last_stats()["flagged_solids"] >= 1
    check_parity(case, o, ref, ctx, got)


def test_replay_seed_is_the_globally_nearest_cell_also_for_centres_outside_the_mesh():
    """the reference seeds the fill at findNearestCell(centre) over the WHOLE mesh (src/solidcloud.cpp:363-365).  Thin
    rectangles whose centres lie outside this (sub)mesh: the nearest cell is then often not a candidate of the solid, and the
    fill starts from the first member in index order (src/cellenumerator.cpp:52-63) — the kept component must be the oracle's."""
    kept_fewer = 0
    for pos0, e0, pos3, e3 in [((-2.2, 9.1, 0), 31, (12.0, 26.0, 0), 75), ((24.7, 12.3, 0), 143, (3.1, -1.2, 0), 62)]:
        case = cases.case_disconnected()
        S = case["solids"]
        S[0]["pos"] = pos0; S[0]["quat"] = cases.quat_from_euler_xyz_deg((0, 0, e0))     # needle centres outside the mesh,
        S[3]["pos"] = pos3; S[3]["quat"] = cases.quat_from_euler_xyz_deg((0, 0, e3))     # the needles reach in
        o, ref, ctx, got = run_both(case)
        assert ctx.last_stats()["flagged_solids"] >= 1
        check_parity(case, o, ref, ctx, got)
        from oracle.oracle_py import eval_points
        for s_ in (0, 3):                                                              # the fill really dropped a component
            inside, _ = eval_points(case["shapes"], S[s_], case["mesh"].points)
            n_members = int(inside[case["mesh"].cp.reshape(-1, 8)].any(axis=1).sum())
            kept_fewer += int(ref["list_off"][3 * s_ + 3] - ref["list_off"][3 * s_]) < n_members
    assert kept_fewer >= 3
    case = cases.case_disconnected()
    o, ref, ctx, got = run_both(case)
    assert ctx.last_stats()["flagged_solids"] >= 1
    check_parity(case, o, ref, ctx, got)
    assert ref["list_off"][-1] > 20


def test_wrong_no_global_promise_is_caught_and_the_step_run_again():
    """sdfibm_interact_device_solids(may_be_global = 0) with a plane present used to drop the plane silently."""
    import torch

    case = cases.case_c2(with_walls=True)        # 100 circles + 4 wall planes
    o, ref, ctx, got = run_both(case)
    dev = torch.device("cuda", 0)
    nC, nS = case["mesh"].n_cells, len(case["solids"])
    dS = torch.from_numpy(np.ascontiguousarray(case["solids"]).view(np.uint8)).to(dev)
    dU = torch.from_numpy(case["U"]).to(dev)
    f = [torch.zeros(k, dtype=torch.float64, device=dev) for k in (nC, 3 * nC, nC, nC, 6 * nS)]
    ctx.interact_device_solids(dS.data_ptr(), nS, dU.data_ptr(), case["dt"], case["rhof"], *[x.data_ptr() for x in f], may_be_global=False)
    assert np.array_equal(f[3].cpu().numpy(), got["Ct"]) and np.array_equal(f[0].cpu().numpy(), got["As"])
    assert np.array_equal(f[1].cpu().numpy().reshape(nC, 3), got["Fs"]) and np.array_equal(f[2].cpu().numpy(), got["Ts"])
    off, cells = ctx.candidate_lists()
    assert np.array_equal(off, ref["list_off"]) and np.array_equal(cells, ref["list_cells"])     # the planes' cells are there
    ft = f[4].cpu().numpy().reshape(nS, 6)
    assert np.abs(ft - got["FT"]).max() <= 1e-10 * np.abs(got["FT"]).max() and np.abs(ft[-4:]).max() > 0   # the four walls carry load


def test_empty_and_outside_solids():
    """Solids that touch no cell of this (sub)mesh yield empty lists and zero force (cellenumerator.cpp:52-65)."""
    case = cases.case_c4(n=32, n_solids=6, n_side=2)
    case["solids"]["pos"][0] = (-40.0, 5.0, 5.0)
    case["solids"]["pos"][1] = (16.0, 16.0, 90.0)
    o, ref, ctx, got = run_both(case)
    check_parity(case, o, ref, ctx, got)
    assert not got["FT"][:2].any()


def test_slot_overflow_widens_the_records_and_runs_again():
    """More solids touch a cell than there are slot records: the step widens them and runs again (no error, same results)."""
    case = cases.case_c4(n=16, n_solids=5, n_side=1)
    case["solids"]["pos"][:] = (8.0, 8.0, 8.0)
    case["solids"]["pos"][:, 0] += np.arange(5) * 0.37
    o, ref, ctx, got = run_both(case, cell_slots=2)
    check_parity(case, o, ref, ctx, got)
    again = ctx.interact(case["solids"], case["U"], case["dt"], case["rhof"])   # the widened records stay
    assert np.array_equal(again["As"], got["As"]) and np.array_equal(again["Ct"], got["Ct"])


def test_more_candidates_than_tile_slots_switches_to_scan_and_fill():
    """The default binning gives every tile a fixed number of list slots (BIN_FIXED_CAP = 8: no scan, no fill pass).  A tile that 14
    solids reach does not fit: the step is run again with the scan + fill lists (sized exactly), and stays on them."""
    case = cases.case_c4(n=24, n_solids=14, n_side=3)
    rng = np.random.RandomState(11)
    case["solids"]["pos"][:] = (9.0, 10.0, 11.0) + rng.uniform(-2.5, 2.5, size=(14, 3))     # all within one 8 x 8 x 4 tile's reach
    o, ref, ctx, got = run_both(case, cell_slots=16)
    check_parity(case, o, ref, ctx, got)
    assert ctx.last_stats()["bin_entries"] >= 14 * 8       # every solid reaches several tiles; the count is the true total
    again = ctx.interact(case["solids"], case["U"], case["dt"], case["rhof"])
    assert np.array_equal(again["As"], got["As"]) and np.array_equal(again["Ct"], got["Ct"])
    # the same step on a context that never used the fixed slots
    os.environ["SDFIBM_BIN_FIXED"] = "0"
    try:
        _, _, ctx2, got2 = run_both(case, cell_slots=16)
    finally:
        del os.environ["SDFIBM_BIN_FIXED"]
    assert np.array_equal(got2["As"], got["As"]) and np.array_equal(got2["Ct"], got["Ct"]) and np.array_equal(got2["Fs"], got["Fs"])
    assert ctx2.last_stats()["bin_entries"] == ctx.last_stats()["bin_entries"]


def test_collision_parity():
    case = cases.case_c2(with_walls=True)
    rng = np.random.RandomState(11)
    S = case["solids"].copy()
    S["pos"][:100, :2] += rng.uniform(-0.12, 0.12, size=(100, 2))   # create overlaps, keep planes in place
    o = Oracle(case["mesh"], True)
    ctx = Context(0)
    ctx.set_mesh(case["mesh"], True)
    ctx.set_shapes(case["shapes"])
    for delta in (-2.0, 0.3, 0.45):
        pr, fr = o.collide(case["shapes"], S, delta)
        pg, fg = ctx.collide(S, delta)
        assert np.array_equal(pr, pg), delta
        assert np.abs(fr - fg).max() <= REL_FORCE * max(1.0, np.abs(fr).max())
    assert len(pr) > 0 and np.abs(fr).max() > 0
    # 3-D spheres + a plane
    c3 = cases.case_mixed3d()
    pr, fr = o.__class__(c3["mesh"], False).collide(c3["shapes"], c3["solids"], 7.0)
    ctx3 = Context(0)
    ctx3.set_mesh(c3["mesh"], False)
    ctx3.set_shapes(c3["shapes"])
    pg, fg = ctx3.collide(c3["solids"], 7.0)
    assert np.array_equal(pr, pg) and len(pr) > 0
    assert np.abs(fr - fg).max() <= REL_FORCE * max(1.0, np.abs(fr).max())


def test_sharded_blocks_on_gpu():
    """N>1 data path on the GPU: every rank's block (solids replicated, many of them outside or straddling the
    block) against the oracle on the same block, and the sum of the per-rank partial force/torque against the
    oracle's (what the NCCL all-reduce delivers; the collective itself is covered by tests/test_multirank_cpu.py
    on gloo and by bench.py --gpus N on NCCL)."""
    world = 2
    total_gpu, total_ref = 0.0, 0.0
    for rank in range(world):
        case = cases.case_c5_block(rank, world, n=32, n_solids=24, n_side=3)
        o, ref, ctx, got = run_both(case)
        check_parity(case, o, ref, ctx, got)
        total_gpu = total_gpu + got["FT"]
        total_ref = total_ref + ref["FT"]
    assert np.abs(total_gpu - total_ref).max() <= REL_FORCE * np.abs(total_ref).max()


def test_full_size_properties_c4():
    """C4 at BASELINE size (256^3, 10^4 spheres): size-independent properties instead of an oracle run.
    - pair counts: every sphere r=5 touches 600..1000 cells; totals match the survey's estimate (~7.9e6)
    - As in [0,1]; Ts >= As; Ct in {0,2,3} or >= 4 and As == 1 where Ct >= 4
    - sum(As*V) ~ N * 4/3 pi r^3 within the algorithm's known -2..-3 % bias
    - the solid sub-range [0,40) of the same pack reproduces the oracle bit-for-bit in its lists
    - idempotence: a second interact gives identical fields."""
    case = cases.case_c4()
    ctx = Context(0)
    ctx.set_mesh(case["mesh"], False)
    ctx.set_shapes(case["shapes"])
    got = ctx.interact(case["solids"], case["U"], case["dt"], case["rhof"])
    counts = ctx.candidate_counts()
    assert 7.5e6 < sum(counts) < 8.3e6
    off, cells = ctx.candidate_lists()
    per_solid = off[3::3] - off[0:-1:3]
    assert per_solid.min() >= 600 and per_solid.max() <= 1000
    for s in range(0, 10000, 997):   # every list segment is ascending (std::set order)
        for t in range(3):
            seg = cells[off[3 * s + t]:off[3 * s + t + 1]]
            assert (np.diff(seg) > 0).all()
    As, Ts, Ct = got["As"], got["Ts"], got["Ct"]
    assert As.min() >= 0.0 and As.max() <= 1.0 and (Ts >= As - 1e-15).all()
    assert set(np.unique(Ct[Ct < 4]).tolist()) <= {0.0, 2.0, 3.0}
    assert (As[Ct >= 4] == 1.0).all()
    vol = float((As * case["mesh"].V).sum())
    exact = 10000 * 4.0 / 3.0 * np.pi * 125.0
    assert -0.04 < vol / exact - 1.0 < -0.01
    again = ctx.interact(case["solids"], case["U"], case["dt"], case["rhof"])
    for k in ("As", "Fs", "Ts", "Ct"):
        assert np.array_equal(got[k], again[k])
    assert np.abs(got["FT"] - again["FT"]).max() <= 1e-10 * np.abs(got["FT"]).max()
    # oracle on a bounded sample of the same workload: 500 evenly spaced solids (each evaluated alone on the whole mesh)
    idx = np.linspace(0, len(case["solids"]) - 1, 500).astype(int)
    sub = np.ascontiguousarray(case["solids"][idx])
    ref = Oracle(case["mesh"], False).interact(case["shapes"], sub, case["U"], case["dt"], case["rhof"], faithful=False)
    for j, s_ in enumerate(idx):
        for t in range(3):
            assert np.array_equal(cells[off[3 * s_ + t]: off[3 * s_ + t + 1]], ref["list_cells"][ref["list_off"][3 * j + t]: ref["list_off"][3 * j + t + 1]])
    owners = np.bincount(cells, minlength=case["mesh"].n_cells)
    mine = np.unique(ref["list_cells"])
    mine = mine[owners[mine] == 1]                      # cells no other solid touches: the sample's fields are the full run's
    assert len(mine) > 300_000
    assert_fields_close({k: got[k][mine] for k in ("As", "Ts", "Fs")}, {k: ref[k][mine] for k in ("As", "Ts", "Fs")})
    scale = _term_scale(dict(case, solids=sub), ref["list_off"], ref["list_cells"])
    err = np.abs(got["FT"][idx] - ref["FT"])
    assert (err <= REL_FORCE * np.maximum(scale, np.abs(ref["FT"])) + 1e-300).all(), err.max()


def test_empty_cloud_resets_the_fields():
    """No solids: interact only zeroes As/Fs/Ts/Ct (reference src/solidcloud.cpp:438-441); fixInternal changes nothing."""
    case = cases.case_c4(n=16, n_solids=2, n_side=1)
    ctx = Context(0)
    ctx.set_mesh(case["mesh"], False)
    ctx.set_shapes(case["shapes"])
    first = ctx.interact(case["solids"], case["U"], case["dt"], case["rhof"])
    assert first["As"].max() > 0
    out = ctx.interact(case["solids"][:0], case["U"], case["dt"], case["rhof"])
    for k in ("As", "Fs", "Ts", "Ct"):
        assert not out[k].any(), k
    assert ctx.candidate_counts() == [0, 0, 0]
    assert np.array_equal(ctx.fix_internal(case["solids"][:0], case["U"]), case["U"])


def test_mean_field_sampler():
    """SolidCloud::calcMeanField (src/solidcloud.cpp:315-359): sum(alpha V U) / sum(alpha V) over a substitute shape per
    solid, against the same sums formed from the oracle's per-solid alpha (one solid at a time, so As == alpha)."""
    case = cases.case_mixed3d(n=24, n_solids=6)
    S = case["solids"][:6].copy()
    S["shape"] = [0, 1, 2, 3, 4, 0]          # substitute shapes: Sphere, Ellipsoid, Box, Sphere+com, Box+com, Sphere
    ctx = Context(0, cell_slots=8)
    ctx.set_mesh(case["mesh"], False)
    ctx.set_shapes(case["shapes"])
    before = ctx.interact(case["solids"], case["U"], case["dt"], case["rhof"])
    mean, den = ctx.mean_field(S, case["U"])
    o = Oracle(case["mesh"], False)
    V = case["mesh"].V
    for i in range(len(S)):
        one = S[i:i + 1].copy()
        one["vel"] = 0
        one["omega"] = 0
        a = o.interact(case["shapes"], one, case["U"], 1.0, 1.0)["Ts"]       # unclamped alpha of this solid alone
        if a.sum() == 0:
            continue
        ref_den = float((a * V).sum())
        ref_mean = (a * V) @ case["U"] / ref_den
        assert abs(den[i] - ref_den) <= 1e-10 * ref_den
        assert np.abs(mean[i] - ref_mean).max() <= 1e-10 * max(1.0, np.abs(ref_mean).max())
    # the sampler must not disturb the coupling state: fixInternal still uses the Ct of the last interact
    assert np.array_equal(ctx.fix_internal(case["solids"], case["U"]),
                          o.fix_internal(case["shapes"], case["solids"], before["Ct"], case["U"]))

"""Row f2: the step either side of interact on the device (reference src/main.cpp:70-77) and the compact touched-cell records."""
import numpy as np
import pytest

from sdfibm_b200 import cases


def _ctx(case):
    from sdfibm_b200.context import Context
    ctx = Context(0)
    ctx.set_mesh(case["mesh"], case["two_d"])
    ctx.set_shapes(case["shapes"])
    return ctx


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c4_small", "mixed3d", "c2_walls"])
def test_apply_forcing_on_the_device_is_main_cpp_70_77_bit_for_bit(name):
    import torch

    case = {"c4_small": lambda: cases.case_c4(n=32, n_solids=8, n_side=2), "mixed3d": cases.case_mixed3d,
            "c2_walls": lambda: cases.case_c2(with_walls=True)}[name]()
    ctx = _ctx(case)
    nC, nS = case["mesh"].n_cells, len(case["solids"])
    dev = torch.device("cuda", 0)
    rng = np.random.RandomState(3)
    T0 = 300.0 + rng.standard_normal(nC)
    dU = torch.from_numpy(case["U"]).to(dev)
    dT = torch.from_numpy(T0).to(dev)
    f = [torch.empty(n, dtype=torch.float64, device=dev) for n in (nC, 3 * nC, nC, nC, 6 * nS)]
    ctx.interact_device(case["solids"], dU.data_ptr(), case["dt"], case["rhof"], *[x.data_ptr() for x in f])
    As, Fs, Ts, Ct = (x.cpu().numpy() for x in f[:4])
    Fs = Fs.reshape(nC, 3)
    ctx.apply_forcing_device(dU.data_ptr(), dT.data_ptr(), case["dt"])
    ctx_sync = torch.cuda.synchronize()
    U1, T1 = dU.cpu().numpy().reshape(nC, 3), dT.cpu().numpy()
    assert np.array_equal(U1, case["U"] - Fs * case["dt"])                       # U = U - Fs*dt
    assert np.array_equal(T1, (1.0 - As) * T0 + Ts)                              # T = (1.0 - As)*T + Ts
    # the same fields as the host-buffer entry gives
    host = ctx.interact(case["solids"], case["U"], case["dt"], case["rhof"])
    assert np.array_equal(host["As"], As) and np.array_equal(host["Fs"], Fs) and np.array_equal(host["Ct"], Ct)
    # compact records: exactly the cells where any field is non-zero are a subset of the touched cells, values equal
    tc = ctx.touched_cells()
    dense_nz = np.nonzero((host["As"] != 0) | (host["Ts"] != 0) | (host["Ct"] != 0) | (np.abs(host["Fs"]).sum(axis=1) != 0))[0]
    assert len(np.unique(tc["cells"])) == len(tc["cells"]) and np.isin(dense_nz, tc["cells"]).all()
    for k in ("As", "Fs", "Ts", "Ct"):
        assert np.array_equal(tc[k], host[k][tc["cells"]])
    off, cells = ctx.candidate_lists()
    assert np.isin(cells, tc["cells"]).all()
    ctx.close()


@pytest.mark.gpu
def test_resident_entries_refuse_use_before_interact():
    from sdfibm_b200 import capi
    case = cases.case_c4(n=16, n_solids=1, n_side=1)
    ctx = _ctx(case)
    with pytest.raises(capi.SdfibmError):
        ctx.touched_cells()
    with pytest.raises(capi.SdfibmError):
        ctx.apply_forcing_device(1, None, 0.1)
    ctx.close()

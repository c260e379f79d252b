"""The library's own cross-rank exchange (sdfibm_comm_init: NCCL all-gather of the replicated solid slices + ONE all-reduce of the
per-solid force/torque, reference src/solidcloud.cpp:427-431) on two GPUs, one process per GPU, against the single-block run."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
def test_two_rank_allreduce_and_solid_gather_inside_the_library():
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    from oracle.oracle_py import Oracle
    from sdfibm_b200 import cases

    with tempfile.TemporaryDirectory() as d:
        procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "comm_worker.py"), str(r), "2", d],
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
        outs = [p.communicate(timeout=240)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(outs)
        r0, r1 = (np.load(os.path.join(d, f"rank{r}.npz")) for r in range(2))
    # every rank holds the same, summed array; it is the sum of the ranks' partial sums
    assert np.array_equal(r0["FT"], r1["FT"])
    want = r0["FT_partial"] + r1["FT_partial"]
    scale = np.abs(want).max()
    assert np.abs(r0["FT"] - want).max() <= 1e-13 * scale
    assert np.abs(r0["FT_again"] - want).max() <= 1e-13 * scale and np.array_equal(r0["FT_again"], r1["FT_again"])
    # the device-resident entry (split step: all-reduce alongside the certificate pass, retry flags in a second all-reduce; its
    # first step included rank 1's queue overflow and re-run) gives the same sums on both ranks, step after step
    for k in ("FT_dev0", "FT_dev1", "FT_dev2"):
        assert np.array_equal(r0[k], r1[k]), k
        assert np.abs(r0[k] - want).max() <= 1e-13 * scale, k
    assert np.array_equal(r0["As_dev"], r0["As"]) and np.array_equal(r1["As_dev"], r1["As"])
    # the gathered solid records gave the same fields as the full upload
    assert np.array_equal(r0["As"], r0["As_partial"]) and np.array_equal(r1["As"], r1["As_partial"])
    # ... and the sum is the single-block answer (oracle on the whole mesh)
    whole = cases.case_c5_block(0, 1, n=48, n_solids=60, n_side=4)
    ref = Oracle(whole["mesh"], False).interact(whole["shapes"], whole["solids"], whole["U"], whole["dt"], whole["rhof"])
    assert np.abs(r0["FT"] - ref["FT"]).max() <= 1e-10 * np.abs(ref["FT"]).max()
    assert int(r0["off"][-1]) + int(r1["off"][-1]) == int(ref["list_off"][-1])


@pytest.mark.gpu
def test_comm_entry_points_refuse_misuse():
    from sdfibm_b200 import capi
    from sdfibm_b200.context import Context

    ctx = Context(0)
    with pytest.raises(capi.SdfibmError):
        ctx.allreduce_force_torque(0, 4)            # no communicator yet (and a null pointer)
    uid = Context.comm_unique_id()
    assert len(uid) == 128 and any(uid)
    ctx.comm_init(uid, 0, 1)                         # a one-rank communicator is legal; the step runs without collectives
    with pytest.raises(capi.SdfibmError):
        ctx.comm_init(uid, 0, 1)
    case = cases_small()
    ctx.set_mesh(case["mesh"], case["two_d"])
    ctx.set_shapes(case["shapes"])
    a = ctx.interact(case["solids"], case["U"], case["dt"], case["rhof"])
    ctx.comm_destroy()
    b = ctx.interact(case["solids"], case["U"], case["dt"], case["rhof"])
    assert np.array_equal(a["As"], b["As"]) and np.abs(a["FT"] - b["FT"]).max() <= 1e-12 * np.abs(b["FT"]).max()
    ctx.close()


def cases_small():
    from sdfibm_b200 import cases
    return cases.case_c4(n=24, n_solids=8, n_side=2)

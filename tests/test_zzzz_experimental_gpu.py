"""Opt-in experimental kernel variants (not on the default path; this file sorts last on purpose: a faulting experimental kernel must not poison the CUDA context of any other test).

SDFIBM_SYNTH_FACES=1: k_heavy_hex<., true> forms the face centre / area vector of box cells from the staged vertices instead of
fetching face records.  Written at the end of round 1 with no GPU time left: compiled, never run.  The test is
non-strict xfail — XPASS in the log means the variant meets the parity bars and is ready to be measured."""
import os

import numpy as np
import pytest

from sdfibm_b200 import cases

pytestmark = pytest.mark.gpu


@pytest.mark.xfail(reason="experimental variant, never run on a GPU yet", strict=False)
@pytest.mark.parametrize("variant", ["1", "2"])       # 1: 5 CTAs / SM (90 registers), 2: 6 CTAs / SM (80 registers)
@pytest.mark.parametrize("name", ["c4_small", "c1"])
def test_synthesised_faces_variant_meets_the_parity_bars(name, variant):
    from test_gpu_parity import check_parity, run_both

    case = cases.case_c4(n=48, n_solids=50, n_side=4) if name == "c4_small" else cases.case_c1()
    old = os.environ.get("SDFIBM_SYNTH_FACES")
    os.environ["SDFIBM_SYNTH_FACES"] = variant    # read by sdfibm_create
    try:
        o, ref, ctx, got = run_both(case)
    finally:
        if old is None:
            del os.environ["SDFIBM_SYNTH_FACES"]
        else:
            os.environ["SDFIBM_SYNTH_FACES"] = old
    check_parity(case, o, ref, ctx, got)
    assert np.isfinite(got["As"]).all() and sum(ctx.candidate_counts()) > 0

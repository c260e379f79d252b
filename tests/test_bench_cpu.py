"""bench.py pieces that need no GPU: the reference arm (oracle port on the host cores) and its solid sampling."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    return lines


def test_reference_arm_prints_the_contract_line():
    (line,) = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--cells-per-side", "32"])
    d = json.loads(line)
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["unit"] == "cell-solid updates/s" and d["higher_is_better"] is True and d["value"] > 0
    from oracle import ref_py

    assert d["cpu_baseline"]["kind"] == ("reference" if ref_py.available() else "port")   # oracle/_ref: the reference's own compiled code
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("C4")


def test_reference_arm_under_torchrun_only_rank_zero_works():
    env = {"WORLD_SIZE": "2", "LOCAL_RANK": "1", "RANK": "1"}
    assert _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--cells-per-side", "32"], env) == []
    env = {"WORLD_SIZE": "2", "LOCAL_RANK": "0", "RANK": "0"}
    (line,) = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--cells-per-side", "32"], env)
    d = json.loads(line)
    assert d["n_gpus"] == 2 and d["config"]["workload"].startswith("C5") and d["value"] > 0


def test_representative_solids_keep_the_touching_share():
    import bench
    from sdfibm_b200 import cases

    case = cases.case_c5_block(0, 2, n=32, n_solids=24, n_side=3)       # rank 0 of 2: about half the solids touch its block
    m, S = case["mesh"], case["solids"]
    idx = bench.representative_solids(case, 8)
    assert len(set(idx.tolist())) == len(idx) == 8
    inside = np.all((S["pos"][idx] >= m.bounds_min - 6) & (S["pos"][idx] <= m.bounds_max + 6), axis=1)
    assert 0 < inside.sum() < 8                                           # both kinds are represented
    assert len(bench.representative_solids(case, 1)) == 1

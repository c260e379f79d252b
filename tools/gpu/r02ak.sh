#!/bin/bash
# round 2, call ak: GPU suite + the contract bench line on the final tree (after the tail refactor of the split step)
set -u
R=r02ak
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -rxX > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
tail -4 gpurun_out/${R}_pytest_gpu.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${R}_smoke.log
timeout 900 python bench.py > gpurun_out/${R}_bench_default.json 2> gpurun_out/${R}_bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02ak_bench_default.json") if l.startswith("{")][-1])
k = d.get("kernel_ms", {}); c = d.get("parity_check") or {}
print("ms/step %.4g" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], {a[:10]: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "e2e", d["e2e"]["ms_per_step"], "ok" if c.get("lists_equal") and c.get("Ct_equal") else c, "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], d["clocks"])
PY

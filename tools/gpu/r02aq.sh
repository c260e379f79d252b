#!/bin/bash
# round 2, call aq: last check of the shipped binaries: GPU suite, smoke, default bench line
set -u
R=r02aq
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${R}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/${R}_bench_default.json 2> gpurun_out/${R}_bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02aq_bench_default.json") if l.startswith("{")][-1])
k = d.get("kernel_ms", {}); c = d.get("parity_check") or {}
print("ms/step %.4g" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], {a[:10]: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "e2e", d["e2e"]["ms_per_step"], "ok" if c.get("lists_equal") and c.get("Ct_equal") else c)
PY

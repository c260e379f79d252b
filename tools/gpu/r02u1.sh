#!/bin/bash
# round 2, call u1: C5 (512^3, 1e5 solids) on ONE GPU: the same-workload base of the N = 8 line of call u8; then the GPU suite on the current tree
set -u
R=r02u1
mkdir -p gpurun_out
timeout 1500 python bench.py --workload c5 --gpus 1 --steps 5 --warmup 3 --no-cpu --no-e2e --no-check > gpurun_out/${R}_bench_c5_n1.json 2> gpurun_out/${R}_bench_c5_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/${R}_bench_c5_n1.err | cut -c1-300
timeout 300 python bench.py --workload c3tc --steps 50 --warmup 5 --no-cpu > gpurun_out/${R}_bench_c3tc.json 2> gpurun_out/${R}_bench_c3tc.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02u1_bench*.json")):
    d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    k = d.get("kernel_ms", {})
    print(f[-14:], "ms/step %.4g" % d["ms_per_step"], "value %.4g" % d["value"], {a[:12]: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "e2e", (d.get("e2e") or {}).get("ms_per_step"), (d.get("parity_check") or {}).get("lists_equal"))
PY
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
tail -5 gpurun_out/${R}_pytest_gpu.log

#!/bin/bash
# round 2, call an: the shape partition as a compile-time variant (only for tables mixing spheres with other shapes): GPU suite + C4 / C5 timing
set -u
R=r02an
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
tail -4 gpurun_out/${R}_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err
timeout 300 python bench.py --workload c5 --cells-per-side 256 --steps 8 --warmup 3 --no-cpu > gpurun_out/${R}_bench_c5_256.json 2> gpurun_out/${R}_bench_c5_256.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02an_bench*.json")):
    d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    k = d.get("kernel_ms", {}); c = d.get("parity_check") or {}
    print(f.split("/")[-1][6:-5], "ms/step %.4g" % d["ms_per_step"], "frac %.3f" % (d.get("roofline") or {}).get("frac"),
          {a[:10]: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "e2e", (d.get("e2e") or {}).get("ms_per_step"), "ok" if c.get("lists_equal") and c.get("Ct_equal") else c)
PY

#!/bin/bash
set -u
R=r02n
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -rxX > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
tail -40 gpurun_out/${R}_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err; echo "bench rc=$?"
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02n_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as ex:
        print(f, "no line:", ex); continue
    k = d.get("kernel_ms", {})
    print(f.split("/")[-1], "ms/step %.4g" % d["ms_per_step"], "frac %.3f" % (d.get("roofline") or {}).get("frac"), {a: round(b, 4) for a, b in k.items() if isinstance(b, float)})
PY

#!/bin/bash
# fused pipeline: GPU tests, then C4 bench A/B (fused + box2 | unfused + box2 | unfused + box v1)
set -u
R=r02e
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rxX -x > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
tail -25 gpurun_out/${R}_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err; echo "bench rc=$?"
SDFIBM_FUSED=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-check > gpurun_out/${R}_bench_c4_unfused.json 2> gpurun_out/${R}_bench_c4_unfused.err
SDFIBM_FUSED=0 SDFIBM_BOX_V=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-check > gpurun_out/${R}_bench_c4_v1.json 2> gpurun_out/${R}_bench_c4_v1.err
timeout 600 python bench.py --workload c5 --cells-per-side 256 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench_c5s.json 2> gpurun_out/${R}_bench_c5s.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02e_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as ex:
        print(f, "no line:", ex); continue
    k = d.get("kernel_ms", {})
    print(f.split("/")[-1], "ms/step %.4g" % d["ms_per_step"], "frac", (d.get("roofline") or {}).get("frac"),
          {a: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "e2e", (d.get("e2e") or {}).get("ms_per_step"), "check", d.get("parity_check"))
PY
tail -5 gpurun_out/${R}_bench_c4.err

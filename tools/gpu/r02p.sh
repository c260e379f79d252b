#!/bin/bash
# round 2, call p: full GPU suite on the committed tree + launch list + ncu --set full of the three interact kernels
set -u
R=r02p
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rxX > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${R}_pytest_gpu.log | head -30
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-check > /dev/null 2>&1; echo "ncu launches rc=$?"
NCU_SKIP=3 NCU_COUNT=3 bash tools/gpu/ncu_full.sh ${R} "^(k_classify|k_heavy_box|k_final)$"
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02p_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as ex:
        print(f, "no line:", ex); continue
    k = d.get("kernel_ms", {})
    print(f.split("/")[-1], "ms/step %.4g" % d["ms_per_step"], "frac %.3f" % (d.get("roofline") or {}).get("frac"),
          {a: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "e2e", (d.get("e2e") or {}).get("ms_per_step"))
PY

#!/bin/bash
# round 2, call t: full GPU suite on the current tree (k_classify4, vectorised k_fix_internal, persistent collision buffers), aux lines, ncu of k_classify4
set -u
R=r02t
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
tail -5 gpurun_out/${R}_pytest_gpu.log
timeout 600 python tools/bench_aux.py > gpurun_out/${R}_aux.jsonl 2> gpurun_out/${R}_aux.err; echo "aux rc=$?"; tail -3 gpurun_out/${R}_aux.err
cut -c1-600 gpurun_out/${R}_aux.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err; echo "bench rc=$?"
NCU_SKIP=3 NCU_COUNT=3 bash tools/gpu/ncu_full.sh ${R} "^(k_classify4|k_heavy_box|k_final)$"
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02t_bench*.json")):
    d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    k = d.get("kernel_ms", {})
    print(f.split("/")[-1][5:-5], "ms/step %.4g" % d["ms_per_step"], "frac %.3f" % (d.get("roofline") or {}).get("frac"),
          {a[:10]: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "e2e", (d.get("e2e") or {}).get("ms_per_step"), d.get("parity_check"))
PY

#!/bin/bash
# round 2, call ad: k_classify4 with out-of-line corner refinement (C5 shape tables): parity + A/B against the one-position kernel
set -u
R=r02ad
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
tail -4 gpurun_out/${R}_pytest_gpu.log
run5() { local name=$1; shift
  env "$@" timeout 300 python bench.py --workload c5 --cells-per-side 256 --steps 8 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench5_${name}.json 2> gpurun_out/${R}_bench5_${name}.err
}
run5 cls4 X=1
run5 cls1 SDFIBM_CLASSIFY4=0
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02ad_bench*.json")):
    d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    k = d.get("kernel_ms", {}); c = d.get("parity_check") or {}
    print(f.split("/")[-1][6:-5], "ms/step %.4g" % d["ms_per_step"], "frac %.3f" % (d.get("roofline") or {}).get("frac"),
          {a[:10]: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "ok" if c.get("lists_equal") and c.get("Ct_equal") else c)
PY

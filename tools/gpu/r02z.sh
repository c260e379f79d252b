#!/bin/bash
# round 2, call z: k_final_c (touched cells compacted per warp) parity + A/B
set -u
R=r02z
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
tail -5 gpurun_out/${R}_pytest_gpu.log
run() { local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench_${name}.json 2> gpurun_out/${R}_bench_${name}.err
}
run5() { local name=$1; shift
  env "$@" timeout 300 python bench.py --workload c5 --cells-per-side 256 --steps 8 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench5_${name}.json 2> gpurun_out/${R}_bench5_${name}.err
}
run fc2 X=1; run5 fc2 X=1
run fc0 SDFIBM_FINAL_C=0
for v in fc1 fc4 fc8; do run $v SDFIBM_B200_LIB=build/variants/$v.so; done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02z_bench*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as ex:
        print(f, "no line:", ex); continue
    k = d.get("kernel_ms", {})
    c = d.get("parity_check") or {}
    print(f.split("/")[-1][5:-5], "ms/step %.4g" % d["ms_per_step"], "frac %.3f" % (d.get("roofline") or {}).get("frac"),
          {a[:10]: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "ok" if c.get("lists_equal") and c.get("Ct_equal") else c)
PY

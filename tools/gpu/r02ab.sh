#!/bin/bash
# round 2, call ab: k_heavy_box resident-CTA / face-loop-unroll sweep after the edge compaction
set -u
R=r02ab
mkdir -p gpurun_out
run() { local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench_${name}.json 2> gpurun_out/${R}_bench_${name}.err
}
run base X=1
for v in ctas5 ctas7 ctas8 fun1 fun3 fun6; do run $v SDFIBM_B200_LIB=build/variants/$v.so; done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02ab_bench*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as ex:
        print(f, "no line:", ex); continue
    k = d.get("kernel_ms", {}); c = d.get("parity_check") or {}
    print(f.split("/")[-1][6:-5], "ms/step %.4g" % d["ms_per_step"], "frac %.3f" % (d.get("roofline") or {}).get("frac"),
          {a[:10]: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "ok" if c.get("lists_equal") and c.get("Ct_equal") else c)
PY

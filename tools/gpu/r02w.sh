#!/bin/bash
# round 2, call w: CTA-size sweep of k_final / k_final_pf (C4) and of the one-position k_classify (C5 shape tables)
set -u
R=r02w
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench_${name}.json 2> gpurun_out/${R}_bench_${name}.err
}
run5() { local name=$1; shift
  env "$@" timeout 300 python bench.py --workload c5 --cells-per-side 256 --steps 8 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench5_${name}.json 2> gpurun_out/${R}_bench5_${name}.err
}
for v in fnt128 fnt64 fnt32; do run $v SDFIBM_FINAL_PF=0 SDFIBM_B200_LIB=build/variants/$v.so; done
for v in pf128 pf64 pf32; do run $v SDFIBM_B200_LIB=build/variants/$v.so; done
run5 base SDFIBM_FINAL_PF=0
for v in cls128 cls64; do run5 $v SDFIBM_FINAL_PF=0 SDFIBM_B200_LIB=build/variants/$v.so; done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02w_bench*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as ex:
        print(f, "no line:", ex); continue
    k = d.get("kernel_ms", {})
    c = d.get("parity_check") or {}
    print(f.split("/")[-1][5:-5], "ms/step %.4g" % d["ms_per_step"], "frac %.3f" % (d.get("roofline") or {}).get("frac"),
          {a[:10]: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "ok" if c.get("lists_equal") and c.get("Ct_equal") else c)
PY

#!/bin/bash
# round 2, call aj: streaming stores for the four output fields in k_final (A/B)
set -u
R=r02aj
mkdir -p gpurun_out
run() { local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${R}_bench_${name}.json 2> gpurun_out/${R}_bench_${name}.err
}
run base X=1
run stcs SDFIBM_B200_LIB=build/variants/stcs.so
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02aj_bench*.json")):
    d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    k = d.get("kernel_ms", {}); c = d.get("parity_check") or {}
    print(f.split("/")[-1][6:-5], "ms/step %.4g" % d["ms_per_step"], "frac %.3f" % (d.get("roofline") or {}).get("frac"),
          {a[:10]: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "e2e", (d.get("e2e") or {}).get("ms_per_step"), "ok" if c.get("lists_equal") and c.get("Ct_equal") else c)
PY

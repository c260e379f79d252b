#!/bin/bash
# round 2, call s: tile-size sweep with k_classify4 (finer tiles = fewer candidates per cell)
set -u
R=r02s
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench_${name}.json 2> gpurun_out/${R}_bench_${name}.err
}
run t884 X=1
run t444 SDFIBM_TILE=4,4,4
run t844 SDFIBM_TILE=8,4,4
run t882 SDFIBM_TILE=8,8,2
run t842 SDFIBM_TILE=8,4,2
run t484 SDFIBM_TILE=4,8,4
run t448 SDFIBM_TILE=4,4,8
run t888 SDFIBM_TILE=8,8,8
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02s_bench*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as ex:
        print(f, "no line:", ex); continue
    k = d.get("kernel_ms", {})
    c = d.get("parity_check") or {}
    print(f.split("/")[-1][5:-5], "ms/step %.4g" % d["ms_per_step"], "frac %.3f" % (d.get("roofline") or {}).get("frac"),
          {a[:10]: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "ok" if c.get("lists_equal") and c.get("Ct_equal") else c)
PY

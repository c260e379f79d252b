#!/bin/bash
# round 2, call r: k_classify4 with record prefetch, k_final2 (two positions per thread) parity + A/B
set -u
R=r02r
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
tail -5 gpurun_out/${R}_pytest_gpu.log
SDFIBM_FINAL2=1 timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_zy_gpu_vs_reference.py tests/test_zzz_gpu_fuzz.py tests/test_zw_gpu_resident.py -m gpu -q > gpurun_out/${R}_pytest_final2.log 2>&1; echo "pytest final2 rc=$?" | tee -a gpurun_out/${R}_pytest_final2.log
tail -5 gpurun_out/${R}_pytest_final2.log
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench_${name}.json 2> gpurun_out/${R}_bench_${name}.err
}
run base X=1
run final2 SDFIBM_FINAL2=1
for v in f2c3 f2c6; do run $v SDFIBM_FINAL2=1 SDFIBM_B200_LIB=build/variants/$v.so; done
for v in final3 cls_nopf cls_64_12 cls_128_6 cls_128_8; do run $v SDFIBM_B200_LIB=build/variants/$v.so; done
env SDFIBM_FINAL2=1 timeout 300 python bench.py --workload c5 --cells-per-side 256 --steps 8 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench5_final2.json 2> gpurun_out/${R}_bench5_final2.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02r_bench*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as ex:
        print(f, "no line:", ex); continue
    k = d.get("kernel_ms", {})
    c = d.get("parity_check") or {}
    print(f.split("/")[-1][5:-5], "ms/step %.4g" % d["ms_per_step"], "frac %.3f" % (d.get("roofline") or {}).get("frac"),
          {a[:10]: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "ok" if c.get("lists_equal") and c.get("Ct_equal") else c)
PY

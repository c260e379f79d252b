#!/bin/bash
set -u
R=r02j
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rxX -x > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
tail -8 gpurun_out/${R}_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err; echo "bench rc=$?"
for v in c4 c6 c8 f3 fa1 fa3 f3a3 f5; do
  SDFIBM_B200_LIB=build/variants/$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-check > gpurun_out/${R}_bench_c4_$v.json 2> gpurun_out/${R}_bench_c4_$v.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02j_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as ex:
        print(f, "no line:", ex); continue
    k = d.get("kernel_ms", {})
    c = d.get("parity_check") or {}
    print(f.split("/")[-1], "ms/step %.4g" % d["ms_per_step"], "frac %.3f" % (d.get("roofline") or {}).get("frac"),
          {a: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "check", {k_: c[k_] for k_ in ("lists_equal", "max_rel_As", "max_rel_Fs", "Ct_equal", "max_rel_FT") if k_ in c})
PY
tail -5 gpurun_out/${R}_bench_c4.err

#!/bin/bash
set -u
R=r02l
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rxX > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
tail -40 gpurun_out/${R}_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err; echo "bench rc=$?"
SDFIBM_B200_LIB=build/variants/f3.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-check > gpurun_out/${R}_bench_c4_f3.json 2> gpurun_out/${R}_bench_c4_f3.err
timeout 600 python tools/bench_aux.py > gpurun_out/${R}_aux.jsonl 2> gpurun_out/${R}_aux.err; echo "aux rc=$?"; tail -3 gpurun_out/${R}_aux.err
cat gpurun_out/${R}_aux.jsonl
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02l_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as ex:
        print(f, "no line:", ex); continue
    k = d.get("kernel_ms", {})
    c = d.get("parity_check") or {}
    print(f.split("/")[-1], "ms/step %.4g" % d["ms_per_step"], "frac %.3f" % (d.get("roofline") or {}).get("frac"),
          {a: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "\n  e2e", d.get("e2e"), "\n  e2e_host", (d.get("e2e_host_fields") or {}).get("ms_per_step"), "touched", d.get("touched_download"), "\n  check", {k_: c[k_] for k_ in ("lists_equal", "max_rel_As", "max_rel_Fs", "Ct_equal", "max_rel_FT") if k_ in c})
PY
tail -5 gpurun_out/${R}_bench_c4.err

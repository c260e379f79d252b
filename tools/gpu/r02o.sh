#!/bin/bash
set -u
R=r02o
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rxX > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${R}_pytest_gpu.log | head -30
grep -n "Error\|assert " gpurun_out/${R}_pytest_gpu.log | head -40
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err; echo "bench rc=$?"
for w in c1 c2 c3 c3b; do
  timeout 300 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu > gpurun_out/${R}_bench_$w.json 2> gpurun_out/${R}_bench_$w.err
done
timeout 300 python bench.py --workload c5 --cells-per-side 256 --steps 8 --warmup 3 --no-cpu > gpurun_out/${R}_bench_c5_256.json 2> gpurun_out/${R}_bench_c5_256.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02o_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as ex:
        print(f, "no line:", ex); continue
    k = d.get("kernel_ms", {})
    c = d.get("parity_check") or {}
    print(f.split("/")[-1], "ms/step %.4g" % d["ms_per_step"], "frac %.3f" % (d.get("roofline") or {}).get("frac"),
          {a: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "e2e", (d.get("e2e") or {}).get("ms_per_step"), "e2e_host", (d.get("e2e_host_fields") or {}).get("ms_per_step"),
          "check", {k_: c[k_] for k_ in ("lists_equal", "max_rel_As", "Ct_equal", "max_rel_FT") if k_ in c}, "cpu", (d.get("cpu_baseline") or {}).get("value"))
PY

#!/bin/bash
# round 2, call am: queue partitioned by corner path (un-rotated spheres front, all other pairs back): GPU suite + A/B
set -u
R=r02am
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
tail -4 gpurun_out/${R}_pytest_gpu.log
grep -n "Error\|assert " gpurun_out/${R}_pytest_gpu.log | head -20
for m in 1 0; do
  SDFIBM_PART_SHAPES=$m timeout 300 python bench.py --workload c5 --cells-per-side 256 --steps 8 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench_c5_part$m.json 2> gpurun_out/${R}_bench_c5_part$m.err
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err
timeout 600 python tools/gpu_fuzz.py 250 40 5000 > gpurun_out/${R}_fuzz.json 2> gpurun_out/${R}_fuzz.err; echo "fuzz rc=$?"; cut -c1-600 gpurun_out/${R}_fuzz.json
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02am_bench*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as ex:
        print(f, "no line:", ex); continue
    k = d.get("kernel_ms", {}); c = d.get("parity_check") or {}
    print(f.split("/")[-1][6:-5], "ms/step %.4g" % d["ms_per_step"], "frac %.3f" % (d.get("roofline") or {}).get("frac"),
          {a[:10]: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "ok" if c.get("lists_equal") and c.get("Ct_equal") else c)
PY

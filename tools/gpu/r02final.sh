#!/bin/bash
# round 2, evidence call: GPU suite, the contract bench line (C4, with the CPU baseline), the small configurations, C5 at 256^3, the aux
# kernels, the ncu launch list and the ncu --set full capture of the three interact kernels — all on the committed tree
set -u
R=r02g
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rxX > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
tail -6 gpurun_out/${R}_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --cpu-budget 60 > gpurun_out/${R}_bench_reference_arm.json 2> gpurun_out/${R}_bench_reference_arm.err; echo "reference arm rc=$?"
for w in c1 c2 c3 c3b c3tc; do
  timeout 300 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu > gpurun_out/${R}_bench_$w.json 2> gpurun_out/${R}_bench_$w.err
done
timeout 300 python bench.py --workload c5 --cells-per-side 256 --steps 8 --warmup 3 --no-cpu > gpurun_out/${R}_bench_c5_256.json 2> gpurun_out/${R}_bench_c5_256.err
timeout 600 python tools/bench_aux.py > gpurun_out/${R}_aux.jsonl 2> gpurun_out/${R}_aux.err; echo "aux rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-check > /dev/null 2>&1; echo "ncu launches rc=$?"
NCU_SKIP=3 NCU_COUNT=3 bash tools/gpu/ncu_full.sh ${R} "^(k_classify4|k_heavy_box|k_final)$"
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02g_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as ex:
        print(f, "no line:", ex); continue
    if d.get("impl") == "reference":
        print(f.split("/")[-1], "reference arm value %.4g" % d["value"], d["cpu_baseline"].get("distinct_solids")); continue
    k = d.get("kernel_ms", {})
    c = d.get("parity_check") or {}
    print(f.split("/")[-1], "ms/step %.4g" % d["ms_per_step"], "frac %.3f" % (d.get("roofline") or {}).get("frac"),
          {a[:10]: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "e2e", (d.get("e2e") or {}).get("ms_per_step"), "e2e_host", (d.get("e2e_host_fields") or {}).get("ms_per_step"),
          "ok" if c.get("lists_equal") and c.get("Ct_equal") else c, "cpu", (d.get("cpu_baseline") or {}).get("value"))
PY

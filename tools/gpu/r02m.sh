#!/bin/bash
set -u
R=r02m
mkdir -p gpurun_out
timeout 600 python -X faulthandler bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err; echo "bench rc=$?"
tail -12 gpurun_out/${R}_bench_c4.err
timeout 300 python -m pytest tests/test_zw_gpu_resident.py -m gpu -q > gpurun_out/${R}_pytest.log 2>&1; tail -3 gpurun_out/${R}_pytest.log
timeout 600 python tools/bench_aux.py > gpurun_out/${R}_aux.jsonl 2> gpurun_out/${R}_aux.err; echo "aux rc=$?"; tail -3 gpurun_out/${R}_aux.err
cut -c1-400 gpurun_out/${R}_aux.jsonl
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02m_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as ex:
        print(f, "no line:", ex); continue
    k = d.get("kernel_ms", {})
    print(f.split("/")[-1], "ms/step %.4g" % d["ms_per_step"], "frac %.3f" % (d.get("roofline") or {}).get("frac"),
          {a: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "\n  e2e", (d.get("e2e") or {}).get("ms_per_step"), "e2e_host", (d.get("e2e_host_fields") or {}).get("ms_per_step"), "touched ms", (d.get("touched_download") or {}).get("ms"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-check > /dev/null 2>&1; echo "ncu launches rc=$?"
NCU_SKIP=3 NCU_COUNT=3 bash tools/gpu/ncu_full.sh ${R} "k_classify|k_heavy_box|k_final"

#!/bin/bash
# First GPU call of a round (one B200):  gpurun --timeout 1500 -- 'bash tools/first_gpu_call.sh'
# Everything lands in gpurun_out/; nothing here is a bench value if it ran under ncu.
#   1. GPU test suite (incl. test_zy_gpu_vs_reference, test_zzz_gpu_fuzz, the experimental variants as xfail / XPASS)
#   2. the contract bench line (C4) + the reference arm
#   3. the experimental k_heavy_hex variants (SDFIBM_SYNTH_FACES=1 | 2): device-resident step only
#   4. the small BASELINE configurations C1..C3b (ms/step)
#   5. ncu launch list of the bench command, and one full capture of the per-cell kernels
set -u
mkdir -p gpurun_out
R=${ROUND_TAG:-r02}
timeout 900 python -m pytest tests -m gpu -q -rxX > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
tail -15 gpurun_out/${R}_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2>> gpurun_out/${R}_bench_c4.err
for v in 1 2; do
  SDFIBM_SYNTH_FACES=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench_c4_synth$v.json 2> gpurun_out/${R}_bench_c4_synth$v.err
done
for w in c1 c2 c3 c3b; do
  timeout 300 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu > gpurun_out/${R}_bench_$w.json 2> gpurun_out/${R}_bench_$w.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/*_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as ex:
        print(f, "no line:", ex); continue
    k = d.get("kernel_ms", {})
    print(f.split("/")[-1], "ms/step %.4g" % d["ms_per_step"], "value %.4g" % d["value"], "frac", (d.get("roofline") or {}).get("frac"),
          {a: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "e2e", (d.get("e2e") or {}).get("ms_per_step"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > /dev/null 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:"^k_classify$|^k_heavy_hex$|^k_final$|^k_connectivity$" \
    --launch-skip 8 --launch-count 4 -o gpurun_out/${R} python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/${R}_ncu.log 2>&1
ls -la gpurun_out | tail -20

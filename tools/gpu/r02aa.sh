#!/bin/bash
# round 2, call aa: specialised k_classify4 (fmaf, no 2-D selects / clamps) parity + timing; compute-sanitizer memcheck on small cases
set -u
R=r02aa
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
tail -4 gpurun_out/${R}_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err
timeout 300 python bench.py --workload c5 --cells-per-side 256 --steps 8 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench_c5_256.json 2> gpurun_out/${R}_bench_c5_256.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02aa_bench*.json")):
    d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    k = d.get("kernel_ms", {}); c = d.get("parity_check") or {}
    print(f.split("/")[-1][6:-5], "ms/step %.4g" % d["ms_per_step"], "frac %.3f" % (d.get("roofline") or {}).get("frac"),
          {a[:10]: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "ok" if c.get("lists_equal") and c.get("Ct_equal") else c)
PY
# memcheck: the parity cases that exercise every kernel family on small meshes (classify variants, box / hex / general heavy kernels,
# replay, collision, forcing, programs)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest -q -x -m gpu \
   "tests/test_gpu_parity.py" -k "c1 or c2 or mixed or disconnected or collision or graded or empty or scrambled" > gpurun_out/${R}_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/${R}_memcheck.log
tail -6 gpurun_out/${R}_memcheck.log | cut -c1-300
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/${R}_memcheck.log

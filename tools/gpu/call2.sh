set -u
R=${1:-r02b}
timeout 1500 python -m pytest tests -m gpu -x -q -rxX > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/${R}_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err; echo "bench rc=$?"
python - "gpurun_out/${R}_bench_c4.json" <<'PY'
import json,sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("ms/step %.4g"%d["ms_per_step"], d.get("kernel_ms"), "frac", d["roofline"]["frac"])
except Exception as ex: print("no line", ex)
PY
tail -3 gpurun_out/${R}_bench_c4.err

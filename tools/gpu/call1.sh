set -u
R=r02a
python -c "import __graft_entry__ as g; g.build()" 
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err; echo "bench rc=$?"
for v in 1 2; do
  SDFIBM_SYNTH_FACES=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench_c4_synth$v.json 2> gpurun_out/${R}_bench_c4_synth$v.err
done
for w in c1 c2 c3 c3b; do
  timeout 300 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu > gpurun_out/${R}_bench_$w.json 2> gpurun_out/${R}_bench_$w.err
done
tail -c 1500 gpurun_out/${R}_bench_c4.json
for f in gpurun_out/${R}_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("ms/step %.4g"%d["ms_per_step"], d.get("kernel_ms"), "e2e", (d.get("e2e") or {}).get("ms_per_step"))
except Exception as ex: print("no line", ex)
PY
done

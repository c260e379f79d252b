#!/bin/bash
# round 2, call ao: C5 (512^3) on ONE GPU with the final tree: the same-workload base of the 8-GPU line r02_v12
set -u
R=r02ao
mkdir -p gpurun_out
timeout 1500 python bench.py --workload c5 --gpus 1 --steps 5 --warmup 3 --no-cpu --no-e2e --no-check > gpurun_out/${R}_bench_c5_n1.json 2> gpurun_out/${R}_bench_c5_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02ao_bench_c5_n1.json") if l.startswith("{")][-1])
k = d.get("kernel_ms", {})
print("c5 n1 ms/step %.4g" % d["ms_per_step"], "value %.4g" % d["value"], {a[:12]: round(b, 4) for a, b in k.items() if isinstance(b, float)})
PY

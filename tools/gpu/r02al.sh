#!/bin/bash
# round 2, call al (N GPUs, N = $1): C5 on N GPUs with the final tree (no same-workload base: the 1-GPU run is r02_v10_bench_c5_n1.json)
set -u
N=$1
R=r02al
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N --steps 8 --warmup 3 --no-cpu --no-base > gpurun_out/${R}_bench_c5_n$N.json 2> gpurun_out/${R}_bench_c5_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02al_bench_c5_n$N.json") if l.startswith("{")][-1])
k = d.get("kernel_ms", {}); c = d.get("parity_check") or {}
print("N=$N ms/step %.4g" % d["ms_per_step"], "value %.4g" % d["value"], {a[:10]: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "e2e", (d.get("e2e") or {}).get("ms_per_step"), "ok" if c.get("lists_equal") and c.get("Ct_equal") else c)
PY

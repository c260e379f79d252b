#!/bin/bash
# round 2, call u8: C5 on 8 GPUs through the library's own NCCL exchange (no same-workload base here: that is the 1-GPU call r02u1)
set -u
R=r02u8d
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu --no-base > gpurun_out/${R}_bench_c5_n8.json 2> gpurun_out/${R}_bench_c5_n8.err; echo "bench rc=$?"
tail -5 gpurun_out/${R}_bench_c5_n8.err | cut -c1-300
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02u8d_bench_c5_n8.json") if l.startswith("{")][-1])
k = d.get("kernel_ms", {})
print("ms/step %.4g" % d["ms_per_step"], "value %.4g" % d["value"], {a[:12]: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "e2e", (d.get("e2e") or {}).get("ms_per_step"), "e2e_host", (d.get("e2e_host_fields") or {}).get("ms_per_step"))
print(d.get("parity_check"))
PY

#!/bin/bash
# round 2, call ai (8 GPUs): C5 with the split step (all-reduce alongside the certificate pass) vs the unsplit one
set -u
R=r02ai
mkdir -p gpurun_out
for m in 1 0; do
  SDFIBM_COMM_SPLIT=$m timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2956$m bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --no-base --no-e2e --no-check > gpurun_out/${R}_bench_n8_split$m.json 2> gpurun_out/${R}_bench_n8_split$m.err; echo "bench split=$m rc=$?"
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02ai_bench*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as ex:
        print(f, "no line:", ex); continue
    k = d.get("kernel_ms", {}); c = d.get("parity_check") or {}
    print(f.split("/")[-1][6:-5], "ms/step %.4g" % d["ms_per_step"], {a[:10]: round(b, 4) for a, b in k.items() if isinstance(b, float)})
PY

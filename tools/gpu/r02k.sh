#!/bin/bash
# 2 GPUs: the library's NCCL exchange (tests) + the N=2 bench line exactly as the driver launches it
set -u
R=r02k
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zx_gpu_comm.py -m gpu -q -rxXs > gpurun_out/${R}_pytest_comm.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_comm.log
tail -30 gpurun_out/${R}_pytest_comm.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/${R}_bench_n2.json 2> gpurun_out/${R}_bench_n2.err; echo "bench rc=$?"
tail -5 gpurun_out/${R}_bench_n2.err
python - <<'PY'
import json
for f in ["gpurun_out/r02k_bench_n2.json"]:
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as ex:
        print(f, "no line:", ex); continue
    for k in ("ms_per_step", "value", "kernel_ms", "scaling_base", "speedup_same_workload", "pairs_match_base", "parity_check", "e2e"):
        print(k, d.get(k))
PY

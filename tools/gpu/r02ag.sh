#!/bin/bash
# round 2, call ag (2 GPUs): the split multi-GPU step — library comm test, then C5 at 256^3 per rank on 2 GPUs split vs not
set -u
R=r02ag
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_zx_gpu_comm.py -m gpu -q -x > gpurun_out/${R}_pytest_comm.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_comm.log
tail -15 gpurun_out/${R}_pytest_comm.log | cut -c1-300
for m in 1 0; do
  SDFIBM_COMM_SPLIT=$m timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2955$m bench.py --gpus 2 --cells-per-side 256 --steps 10 --warmup 3 --no-cpu --no-base --no-e2e > gpurun_out/${R}_bench_n2_split$m.json 2> gpurun_out/${R}_bench_n2_split$m.err; echo "bench split=$m rc=$?"
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r02ag_bench*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as ex:
        print(f, "no line:", ex); continue
    k = d.get("kernel_ms", {}); c = d.get("parity_check") or {}
    print(f.split("/")[-1][6:-5], "ms/step %.4g" % d["ms_per_step"], {a[:10]: round(b, 4) for a, b in k.items() if isinstance(b, float)}, "ok" if c.get("lists_equal") and c.get("Ct_equal") else c, c.get("allreduce_vs_sum_of_partials_rel"))
PY

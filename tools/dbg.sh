run() { echo "$@"; env "$@" python bench.py --steps 5 --warmup 2 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernel_ms']; print(round(d['ms_per_step'],3), {a:round(b,3) for a,b in k.items() if a.startswith('k_')})"; }
run SDFIBM_CB=128
run SDFIBM_CB=256
run SDFIBM_CB=512

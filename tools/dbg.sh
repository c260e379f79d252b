python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for d in 0 1 4; do echo "DEBUG=$d"; SDFIBM_DEBUG=$d python bench.py --steps 5 --warmup 2 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernel_ms']; print(round(d['ms_per_step'],3), {a:round(b,3) for a,b in k.items()})"; done

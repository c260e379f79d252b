run() { echo "$@"; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernel_ms']; print(round(d['ms_per_step'],3), {a:(round(b,3) if not isinstance(b,dict) else {x:round(y,1) for x,y in b.items()}) for a,b in k.items()})"; }
run X=1

python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run() { python bench.py --steps 20 --warmup 5 --no-cpu --no-map --sensor $1 --top-k $2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$3 $1', {k: round(x,4) for k,x in d['stages_ms'].items()})"; }
run hdl32e 2048 default
run hdl64e 10000 default
BSHOT_SHOT_THREADS=256 run hdl32e 2048 shot256
BSHOT_SHOT_THREADS=256 run hdl64e 10000 shot256
BSHOT_SHOT_THREADS=64 run hdl32e 2048 shot64
BSHOT_SHOT_THREADS=64 run hdl64e 10000 shot64

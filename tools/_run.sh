python -m pytest tests -m gpu -x -q 2>&1 | tail -6
run() { python bench.py --steps 30 --warmup 5 --no-cpu --no-map --no-c3 --sensor $1 --top-k $2 2>gpurun_out/err_$3.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d['stages_ms']; print('%-10s %s seg %.4f topk %.4f normals %.4f shot %.4f match %.4f frame %.4f' % ('$3','$1',s['seg_ratio'],s['topk'],s['normals'],s['shot_bshot'],s['match'],s['frame']))"; }
run hdl32e 2048 default
BSHOT_EXACT_SUMS=1 run hdl32e 2048 exact

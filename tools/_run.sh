run() { python bench.py --steps 30 --warmup 5 --no-cpu --no-map --no-c3 --sensor $1 --top-k $2 2>gpurun_out/err_$3.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d['stages_ms']; print('%-10s %s seg %.4f topk %.4f normals %.4f shot %.4f match %.4f frame %.4f' % ('$3','$1',s['seg_ratio'],s['topk'],s['normals'],s['shot_bshot'],s['match'],s['frame']))"; }
for v in p75 p100; do
  L=$PWD/b-shot-slam_b200/libbshot_b200_$v.so; [ $v = default ] && L=$PWD/b-shot-slam_b200/libbshot_b200.so
  BSHOT_LIB=$L run hdl32e 2048 $v
  BSHOT_LIB=$L run hdl64e 10000 $v
done
for v in p75s p100s; do BSHOT_LIB=$PWD/b-shot-slam_b200/libbshot_b200_$v.so python bench.py --steps 1 --warmup 3 --no-cpu --no-map --no-c3 2>&1 >/dev/null | grep "knn stats" | tail -1; done

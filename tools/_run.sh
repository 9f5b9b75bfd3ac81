python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "--- fp32 interpolation"; python tools/parity_report.py
echo "--- fp64 interpolation"; BSHOT_LIB=$PWD/b-shot-slam_b200/libbshot_b200_fp64.so python tools/parity_report.py
run() { python bench.py --steps 30 --warmup 5 --no-cpu --no-map --no-c3 --sensor $1 --top-k $2 2>gpurun_out/err_$3.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d['stages_ms']; print('%-10s %s seg %.4f normals %.4f shot %.4f match %.4f frame %.4f' % ('$3','$1',s['seg_ratio'],s['normals'],s['shot_bshot'],s['match'],s['frame']))"; }
for v in default fp64; do
  L=$PWD/b-shot-slam_b200/libbshot_b200_$v.so; [ $v = default ] && L=$PWD/b-shot-slam_b200/libbshot_b200.so
  BSHOT_LIB=$L run hdl32e 2048 $v
  BSHOT_LIB=$L run hdl64e 10000 $v
done

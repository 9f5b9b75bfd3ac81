N=$1
python bench.py --steps 3 --warmup 3 --no-cpu --no-c3 --map-steps 5 2>gpurun_out/one.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=1', round(d['ms_per_step'],4), d['map_match']['ms_per_call'], d['map_match']['collective'])"
for mode in peer nccl; do
  BSHOT_EXCHANGE=$mode timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
    bench.py --gpus $N --steps 3 --warmup 3 --map-steps 30 --map-only > gpurun_out/map_r1h_${N}_$mode.json 2> gpurun_out/map_r1h_${N}_$mode.err
  echo "exit $? gpus $N $mode: $(tail -1 gpurun_out/map_r1h_${N}_$mode.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['ms_per_call'], d['collective'][:50])")"
done

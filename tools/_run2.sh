N=$1
for mode in peer nccl; do
  BSHOT_EXCHANGE=$mode timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
    bench.py --gpus $N --steps 3 --warmup 3 --map-steps 40 --map-only > gpurun_out/map_r1f_${N}_$mode.json 2> gpurun_out/map_r1f_${N}_$mode.err
  echo "exit $? gpus $N $mode: $(tail -1 gpurun_out/map_r1f_${N}_$mode.json | cut -c1-420)"
  grep -i "symmetric memory unavailable\|Error\|error" gpurun_out/map_r1f_${N}_$mode.err | head -5
done

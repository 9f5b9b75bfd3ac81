#!/bin/bash
# usage: tools/scale_map.sh N "tiles..."  -- map-only sharded match timing on N GPUs
N=$1; shift
for NT in "$@"; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + NT)) \
    bench.py --gpus $N --steps 3 --warmup 3 --map-steps 30 --map-only 2>/dev/null | tail -1 > gpurun_out/map_${N}_${NT}.json
  echo "gpus $N tiles $NT: $(cut -c1-200 gpurun_out/map_${N}_${NT}.json)"
done

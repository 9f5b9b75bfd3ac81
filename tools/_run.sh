python bench.py --steps 3 --warmup 3 --no-cpu --no-c3 --map-t 2097152 --map-steps 10 --map-only 2>gpurun_out/c5_shard.err | tee gpurun_out/c5_shard.json | cut -c1-400

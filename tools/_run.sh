python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1f_ref.json 2>> gpurun_out/bench_r1f.err; echo "ref exit $?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-c3 --map-steps 2"
$CMD > gpurun_out/plain_r1f.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_r1f.csv $CMD > gpurun_out/ncu_r1f_a.log 2>&1
$CMD > gpurun_out/plain_r1f.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'seg_ratio_kernel|shot_kernel|tk_|normals_kernel|hamming_top2_kernel|merge_top2' -s 20 -c 12 -o gpurun_out/prof_r1f $CMD > gpurun_out/ncu_r1f_b.log 2>&1
tail -n 1 gpurun_out/ncu_r1f_b.log

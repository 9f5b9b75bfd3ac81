python bench.py > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; echo "bench exit $?"; tail -2 gpurun_out/bench_r1g.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1g.json'))
print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, 'e2e', round(d['e2e']['value']), 'c3', round(d['c3']['frame_reference_normals']['ms_per_frame'],3), 'exact', round(d['exact_mode']['ms_per_frame'],3), 'roofline.frac', round(d['roofline']['frac'],4), d['roofline']['ncu'])
PY

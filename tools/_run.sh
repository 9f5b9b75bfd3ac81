python -m pytest tests -m gpu -x -q 2>&1 | tail -1
T0=$(date +%s); python bench.py > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; echo "bench exit $? in $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/bench_r1d.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1d.json'))
print({k:d[k] for k in ['value','ms_per_step']}, d['stages_ms'])
print('e2e', d['e2e']); print('roofline', d['roofline']); print('cpu', d['cpu_baseline']['value'], 'map', d['map_match']['ms_per_call'], d['map_match']['roofline']['frac'])
print('c3 frame', d['c3']['frame_reference_normals'])
for r in d['c3']['radius_sweep_full_normals']: print(r)
PY

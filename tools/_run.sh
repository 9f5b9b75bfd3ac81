python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --no-cpu > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1h.json'))
print({k:d[k] for k in ['value','ms_per_step']}, d['stages_ms'], 'e2e', round(d['e2e']['ms_per_step'],4), 'c3', round(d['c3']['frame_reference_normals']['ms_per_frame'],3), 'exact', round(d['exact_mode']['ms_per_frame'],3))
PY

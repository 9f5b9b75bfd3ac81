python -m pytest tests/test_detector_edge_gpu.py -m gpu -x -q -k reuse 2>&1 | grep -E "^E|assert" | head -20
python - <<'PY'
import sys, numpy as np
sys.path.insert(0,'tests')
from conftest import load_bshot, load_synth
bs, synth = load_bshot(), load_synth()
ctx = bs.Context(0, 131072, 16384, 16384)
scan = synth.make_scan("hdl32e", 4)
ctx.set_cloud(scan)
idx, _, xyz = ctx.detect_keypoints(3000.0, 300, 0, 512)
cached = ctx.compute_normals(0, 3000.0, 300)[:len(idx)]
fresh = ctx.query_normals(xyz, 3000.0, 300)
fresh_b = ctx.query_normals(xyz, 3000.0, 300)
d = np.abs(cached - fresh).max(1)
print("cached vs fresh: exact rows", (d == 0).mean(), "max", d.max(), "n>1e-3", (d > 1e-3).sum(), "fresh vs fresh max", np.abs(fresh - fresh_b).max())
bad = np.argsort(-d)[:5]
for b in bad: print(b, idx[b], cached[b], fresh[b])
PY

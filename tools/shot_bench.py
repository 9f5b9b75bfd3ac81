"""times bshot_compute_lrf (phases A+B) vs bshot_compute_shot (A-D) on the C2 frame"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from conftest import load_bshot, load_synth
bs, synth = load_bshot(), load_synth()
scan = synth.make_scan(sys.argv[1] if len(sys.argv) > 1 else "hdl32e", 0)
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
ctx = bs.Context(0, 131072, 16384, 16384)
ctx.set_cloud(scan)
idx, rat, xyz = ctx.detect_keypoints(3000.0, 300, 0, K)
ctx.compute_normals(0, 3000.0, 300)
for name, fn in (("lrf", lambda: ctx.compute_lrf(3000.0)), ("shot", lambda: ctx.compute_shot(3000.0, want_shot=False))):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(20): r = fn()
    dt = (time.perf_counter() - t0) / 20
    print(name, f"{dt*1e3:.3f} ms", "sum_nn" , r.get("sum_neighbours") if isinstance(r, dict) else int(r[1].sum()))

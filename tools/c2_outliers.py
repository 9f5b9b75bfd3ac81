"""per-frame device times of the C2 loop sequence: the slowest frames with their stage split and detector counters"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench as b
from conftest import load_bshot, load_synth
if __name__ == "__main__":
    import torch
    bs, synth = load_bshot(), load_synth()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    frames = b.make_sequence("hdl32e", n)
    ctx = bs.Context(0, max_points=max(len(f) for f in frames) + 1024, max_keypoints=2048, max_targets=2048)
    p = bs.default_params(top_k=2048)
    d = [torch.from_numpy(f).cuda() for f in frames]
    ctx.enable_timing(True)
    rows = []
    for rep in range(2):
        rows = []
        for i, f in enumerate(frames):
            ctx.process_frame_dev(d[i].data_ptr(), len(f), 12, p)
            st = ctx.stage_times()
            rows.append((st["frame"], i, len(f), {k: round(v, 3) for k, v in st.items()}, ctx.debug_counters()))
    rows.sort(reverse=True)
    for r in rows[:12]:
        print(json.dumps(r))
    print("median", np.median([r[0] for r in rows]))

"""quick per-stage timing of the frame path (C2: HDL-32E K=2048, C3: HDL-64E K=10000); debug aid, not the bench"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_bshot, load_synth  # noqa: E402

bs, synth = load_bshot(), load_synth()
out = {}
only = os.environ.get("BSHOT_QS_ONLY")
for name, sensor, K, mode in (("C2", "hdl32e", 2048, 0), ("C3", "hdl64e", 10000, 0), ("C3_full", "hdl64e", 10000, 1)):
    if only and name != only:
        continue
    frames = [synth.make_scan(sensor, f) for f in range(4)]
    d = [torch.from_numpy(f).cuda() for f in frames]
    ctx = bs.Context(0, max_points=max(len(f) for f in frames) + 1024, max_keypoints=K, max_targets=K)
    p = bs.default_params(top_k=K, normals_mode=mode)
    ctx.enable_timing(True)
    acc = {}
    reps = 12
    for i in range(3 + reps):
        ctx.process_frame_dev(d[i % 4].data_ptr(), len(frames[i % 4]), 12, p)
        if i >= 3:
            for k, v in ctx.stage_times().items():
                acc[k] = acc.get(k, 0.0) + v / reps
    out[name] = {k: round(v, 4) for k, v in acc.items()}
    out[name]["counters"] = ctx.frame_counters()
    out[name]["debug"] = ctx.debug_counters()
    # detector alone, for its own counters
    ctx.set_cloud(frames[0])
    before = ctx.debug_counters()
    ctx.seg_ratio(3000.0, 300, 0)
    after = ctx.debug_counters()
    out[name]["detector_only"] = {k: after[k] - before[k] for k in after}
    ctx.close()
print(json.dumps(out, indent=1))

import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench as b
from conftest import load_bshot, load_synth
if __name__ == "__main__":
    import torch
    bs, synth = load_bshot(), load_synth()
    lo, hi = 184, 192
    frames = [synth.make_scan("hdl32e", f, pos=b.loop_pose(f, 500)[0], yaw_deg=b.loop_pose(f, 500)[1]) for f in range(lo, hi)]
    ctx = bs.Context(0, max_points=max(len(f) for f in frames) + 1024, max_keypoints=2048, max_targets=2048)
    p = bs.default_params(top_k=2048)
    d = [torch.from_numpy(f).cuda() for f in frames]
    st = torch.cuda.ExternalStream(ctx.stream)
    flush = torch.empty((256 << 20) // 4, dtype=torch.float32, device="cuda")
    for timing in (True, False):
        ctx.reset(); ctx.enable_timing(timing)
        for i, f in enumerate(frames):
            with torch.cuda.stream(st):
                flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            ctx.process_frame_dev(d[i].data_ptr(), len(f), 12, p)
            e1.record(st)
            e1.synchronize()
            row = {"frame": lo + i, "n": len(f), "ms": round(e0.elapsed_time(e1), 3)}
            if timing:
                row["stages"] = {k: round(v, 3) for k, v in ctx.stage_times().items()}
                row["dbg"] = ctx.debug_counters()
            print(json.dumps(row))

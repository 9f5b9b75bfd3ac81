"""Matcher kernel crossover: device time of match_shard_dev (left top-2 + reverse pass) per kernel kind and problem size."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from conftest import load_bshot, load_synth


def main():
    import torch
    bs, synth = load_bshot(), load_synth()
    sizes = [(64, 4096), (128, 16384), (256, 256), (600, 600), (600, 5000), (600, 60000), (2048, 2048), (2048, 20000), (4096, 4096),
             (10000, 10000), (10000, 100000), (600, 1 << 20), (2048, 1 << 20), (10000, 1 << 20)]
    ctx = bs.Context(0, max_points=1024, max_keypoints=16384, max_targets=1 << 20)
    st = torch.cuda.ExternalStream(ctx.stream)
    tall = synth.random_descriptors(1 << 20, seed=7)
    out = []
    for nq, nt in sizes:
        ctx.map_reset()
        ctx.map_append(tall[:nt])
        dq = torch.from_numpy(synth.random_descriptors(nq, seed=8).view(np.int64)).cuda()
        cand = torch.empty((nq, 3), dtype=torch.int64, device="cuda")
        row = {"nq": nq, "nt": nt}
        ref = None
        for kind in (0, 1, 2):
            ctx.set_matcher(kind)
            for _ in range(3):
                ctx.match_shard_dev(dq.data_ptr(), nq, 0, 1, cand.data_ptr())
            ctx.sync()
            reps = 20 if nq * nt < 1e9 else 5
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(st):
                e0.record(st)
                for _ in range(reps):
                    ctx.match_shard_dev(dq.data_ptr(), nq, 0, 1, cand.data_ptr())
                e1.record(st)
            e1.synchronize()
            row[f"ms_kind{kind}"] = round(e0.elapsed_time(e1) / reps, 4)
            c = cand.cpu().numpy().copy()
            if ref is None:
                ref = c
            else:
                assert np.array_equal(ref, c), (nq, nt, kind)
        out.append(row)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()

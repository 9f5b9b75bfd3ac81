"""Stand-alone matcher timing (device-resident shard): pairs/s, algorithmic POPC/s vs measured peak."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from conftest import load_bshot, load_synth

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nq", type=int, default=10000)
    ap.add_argument("--nt", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--rq", type=int, default=1)
    a = ap.parse_args()
    import torch
    bs, synth = load_bshot(), load_synth()
    ctx = bs.Context(0, max_points=1024, max_keypoints=max(a.nq, 1024), max_targets=a.nt)
    peak = ctx.popc_peak()
    t = synth.random_descriptors(a.nt, seed=7)
    q = synth.random_descriptors(a.nq, seed=8)
    ctx.map_append(t)
    dq = torch.from_numpy(q.view(np.int64)).cuda()
    cand = torch.empty((a.nq, 3), dtype=torch.int64, device="cuda")
    st = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(3):
        ctx.match_shard_dev(dq.data_ptr(), a.nq, 0, a.rq, cand.data_ptr())
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(a.reps):
            ctx.match_shard_dev(dq.data_ptr(), a.nq, 0, a.rq, cand.data_ptr())
        e1.record(st)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    pairs = a.nq * a.nt
    print(json.dumps(dict(nq=a.nq, nt=a.nt, ms=ms, pairs_per_s=pairs / ms * 1e3, popc_per_s=11 * pairs / ms * 1e3,
                          popc_peak=peak, frac=11 * pairs / ms * 1e3 / peak, target_GBps=a.nt * 48 / ms * 1e-6)))

if __name__ == "__main__":
    main()

"""The six-launch sharded call on ONE rank with a shard of T / 8 descriptors: per-launch device times under
`ncu --metrics gpu__time_duration.sum` show what the call costs besides the shard search."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from conftest import load_bshot, load_sharded, load_synth
import torch
bs, synth, sharded = load_bshot(), load_synth(), load_sharded()
Q, T = int(sys.argv[1]) if len(sys.argv) > 1 else 10000, int(sys.argv[2]) if len(sys.argv) > 2 else 131072
ctx = bs.Context(0, max_points=1024, max_keypoints=Q, max_targets=T)
ctx.set_matcher(int(os.environ.get("KIND", "2")))
ctx.map_append(synth.random_descriptors(T, seed=7))
dq = torch.from_numpy(synth.random_descriptors(Q, seed=8).view(np.int64)).cuda()
m = sharded.DeviceShardedMatcher(ctx, 1, 0, Q, device=torch.device("cuda", 0))
st = torch.cuda.ExternalStream(ctx.stream)
for _ in range(3):
    m.match(dq.data_ptr(), 0)
ctx.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(10):
    m.match(dq.data_ptr(), 0)
e1.record(st)
e1.synchronize()
print("ms per call", e0.elapsed_time(e1) / 10)

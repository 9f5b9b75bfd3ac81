"""in-process multi-rank case of tests/test_sharded_ipc.py (run as a script with CUDA_DEVICE_MAX_CONNECTIONS=32)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_bshot, load_synth  # noqa: E402

bshot, synth = load_bshot(), load_synth()
T, Q, R = 30000, 900, 3
tmap = synth.random_descriptors(T, seed=3, density=60)
q = synth.random_descriptors(Q, seed=4, density=60)
q[:50] = tmap[np.random.default_rng(5).permutation(T)[:50]]
tmap[T - 40:] = tmap[:40]                                   # duplicates in the last shard: the first copy must win
ctxs = [bshot.Context(0, 1024, 1024, T) for _ in range(R)]
whole = bshot.Context(0, 1024, 1024, T)
whole.map_append(tmap)
want = whole.match_map(q, 0)
per = (T + R - 1) // R
for r, c in enumerate(ctxs):
    c.map_append(tmap[r * per:min(T, (r + 1) * per)])
    c.comm_create(r, R, Q)
regions = [c.comm_region()[0] for c in ctxs]
for c in ctxs:
    c.comm_import_ptrs(regions)
dq = torch.from_numpy(q.view(np.int64)).cuda()
outs = [torch.empty((Q, 3), dtype=torch.int64, device="cuda") for _ in ctxs]
torch.cuda.synchronize()
for _ in range(2):                                          # all ranks enqueue, nobody waits on the host in between
    for r, c in enumerate(ctxs):
        c.match_map_sharded_dev(dq.data_ptr(), Q, r * per, outs[r].data_ptr())
for c in ctxs:
    c.sync()
    c.comm_check()
for o in outs:
    got = o.cpu().numpy().view(bshot.CAND_DTYPE).reshape(Q)
    assert np.array_equal(got["k1"], want["k1"]) and np.array_equal(got["k2"], want["k2"])
    assert np.array_equal(got["rq"], want["rq"])
for c in ctxs + [whole]:
    c.close()
print("in-process ranks ok")

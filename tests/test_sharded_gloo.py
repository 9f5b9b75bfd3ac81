"""N > 1 host path on CPU: world_size-2 (and 3) gloo processes shard the map, all-gather the per-rank
candidate records and merge; the result must equal the single-pass oracle (bit-exact, including
cross-shard ties resolved to the lowest global index)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


class OracleBackend:
    """test-only back end: per-shard candidates from the CPU oracle (the GPU back end is GpuBackend)"""

    def __init__(self, oracle, sharded):
        self.oracle, self.sharded, self.shard = oracle, sharded, None

    def set_shard(self, desc):
        self.shard = np.ascontiguousarray(desc)

    def reverse_owned(self, queries, global_base, merged):
        """oracle mirror of bshot_reverse_owned_dev: best query for the winners that live in this shard"""
        rq = np.full(len(queries), -1, np.int32)
        idx = (merged["k1"] & np.uint64(0xFFFFFFFF)).astype(np.int64)
        mine = (merged["k1"] != self.sharded.NONE_KEY) & (idx >= global_base) & (idx < global_base + len(self.shard))
        if mine.any():
            tg = self.shard[idx[mine] - global_base]
            m = self.oracle.match(tg, queries, want_right=False)
            rq[mine] = m["left_idx"]
        return rq

    def match_shard(self, queries, global_base, with_rq=True):
        q = len(queries)
        rec = np.zeros(q, self.sharded.CAND_DTYPE)
        rec["k1"] = rec["k2"] = self.sharded.NONE_KEY
        rec["rq"] = 0xFFFFFFFF
        if len(self.shard):
            m = self.oracle.match(queries, self.shard, want_right=True)
            h1, h2 = m["left_idx"] >= 0, m["left_idx2"] >= 0
            rec["k1"][h1] = (m["left_dist"][h1].astype(np.uint64) << np.uint64(32)) | (m["left_idx"][h1] + global_base).astype(np.uint64)
            rec["k2"][h2] = (m["left_dist2"][h2].astype(np.uint64) << np.uint64(32)) | (m["left_idx2"][h2] + global_base).astype(np.uint64)
            rec["rq"][h1] = m["right_idx"][m["left_idx"][h1]]
        return rec

    def merge(self, gathered):
        return self.sharded.merge_records(gathered)


def _worker(rank, world, port, nq, nt, out_dir):
    import torch
    import torch.distributed as dist
    from conftest import _load, load_oracle, load_synth, PKG
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sharded = _load("bshot_b200_sharded", os.path.join(PKG, "sharded.py"))
    oracle, synth = load_oracle(), load_synth()
    t = synth.random_descriptors(nt, seed=32, density=40)
    q = synth.random_descriptors(nq, seed=31, density=40)
    if nt > 5:
        t[nt - 1] = t[5]      # cross-shard duplicate: the lowest global index must win
        q[3] = t[5]
        q[4] = t[nt - 1]

    def all_gather(local):
        x = torch.from_numpy(local.view(np.int64).copy())
        outs = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(outs, x)
        return np.stack([o.numpy().view(sharded.CAND_DTYPE).reshape(-1) for o in outs])

    def all_reduce_max(local):
        x = torch.from_numpy(local.copy())
        dist.all_reduce(x, op=dist.ReduceOp.MAX)
        return x.numpy()

    sm = sharded.ShardedMap(OracleBackend(oracle, sharded), rank, world, all_gather)
    sm.load(t)
    merged = sm.match(q)
    merged2 = sm.match_sharded_reverse(q, all_reduce_max)     # sharded reverse pass: same records
    assert np.array_equal(merged, merged2), "sharded reverse pass differs from the rq-in-record protocol"
    np.save(os.path.join(out_dir, f"merged_{rank}.npy"), merged)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nq,nt", [(2, 200, 3001), (3, 64, 1000), (2, 50, 1)])
def test_sharded_match_equals_single_pass(tmp_path, oracle, synth, world, nq, nt):
    from conftest import _load, PKG
    sharded = _load("bshot_b200_sharded", os.path.join(PKG, "sharded.py"))
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, nq, nt, str(tmp_path)), nprocs=world, join=True)
    t = synth.random_descriptors(nt, seed=32, density=40)
    q = synth.random_descriptors(nq, seed=31, density=40)
    if nt > 5:
        t[nt - 1] = t[5]
        q[3] = t[5]
        q[4] = t[nt - 1]
    o = oracle.match(q, t, want_right=True)
    for r in range(world):
        m = np.load(os.path.join(str(tmp_path), f"merged_{r}.npy"))
        idx1 = (m["k1"] & np.uint64(0xFFFFFFFF)).astype(np.int64)
        assert np.array_equal(idx1, o["left_idx"])
        assert np.array_equal((m["k1"] >> np.uint64(32)).astype(np.int64), o["left_dist"])
        has2 = m["k2"] != sharded.NONE_KEY
        assert np.array_equal(has2, o["left_idx2"] >= 0)
        assert np.array_equal((m["k2"][has2] & np.uint64(0xFFFFFFFF)).astype(np.int64), o["left_idx2"][has2])
        pairs = sharded.ShardedMap.correspondences(m)
        assert np.array_equal(pairs, oracle.mutual(o["left_idx"], o["right_idx"]))
    if nt > 5:
        assert o["left_idx"][3] == 5 and o["left_idx"][4] == 5


def test_shard_ranges_cover_exactly():
    from conftest import _load, PKG
    sharded = _load("bshot_b200_sharded", os.path.join(PKG, "sharded.py"))
    for total in (0, 1, 7, 8, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [sharded.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))

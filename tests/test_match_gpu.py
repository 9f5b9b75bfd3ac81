"""a10/a11 parity: CUDA Hamming search (through the C ABI) vs the oracle restatement of
src/lidar_odometry.cpp:212-242 + minVect (include/bshot_bits.h:6-20). Bit-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


MATCHERS = {"popc": 0, "tensor-core": 1, "tensor-core-pipelined": 2, "tensor-core-pipelined-smem": 3, "by-size": -1}


@pytest.fixture(scope="module", params=list(MATCHERS))
def mctx(request, bshot):
    """the same checks on every distance-matrix kernel (bshot_set_matcher): XOR + POPC (hamming.cu), tcgen05 int8 dot products
    (hamming_tc.cu), the warp-specialised tensor-core pipeline whose multiply leaves the packed (distance, column) key
    (hamming_tc2.cu; query tiles in tensor memory or in shared memory) and the default choice by problem size"""
    ctx = bshot.Context(0, 131072, 16384, 1 << 21)
    ctx.set_matcher(MATCHERS[request.param])
    yield ctx
    ctx.close()


def _check(ctx, oracle, q, t):
    g = ctx.match(q, t, want_right=True)
    o = oracle.match(q, t, want_right=True)
    for k in ("left_idx", "left_dist", "left_idx2", "left_dist2", "right_idx"):
        assert np.array_equal(g[k], o[k]), k
    pairs, dist = ctx.match_mutual(q, t)
    opairs = oracle.mutual(o["left_idx"], o["right_idx"])
    assert np.array_equal(pairs, opairs)
    assert np.array_equal(dist, o["left_dist"][opairs[:, 0]])


@pytest.mark.parametrize("nq,nt", [(1, 1), (1, 2), (7, 5), (600, 600), (600, 1337), (2048, 2048),
                                   (257, 129), (1025, 4097), (3000, 20000)])
def test_random(mctx, oracle, synth, nq, nt):
    q = synth.random_descriptors(nq, seed=nq)
    t = synth.random_descriptors(nt, seed=1000 + nt)
    _check(mctx, oracle, q, t)


def test_sparse_reference_like(mctx, oracle, synth):
    # <= 40 bits set (what the reference's normals quirk yields): many distance ties
    q = synth.random_descriptors(2048, seed=3, density=33)
    t = synth.random_descriptors(5000, seed=4, density=33)
    _check(mctx, oracle, q, t)


def test_duplicates_and_ties(mctx, oracle, synth):
    rng = np.random.default_rng(11)
    t = synth.random_descriptors(4096, seed=5)
    # planted exact duplicates at several indices: lowest index must win
    t[100] = t[3000]
    t[101] = t[3000]
    t[4095] = t[7]
    q = t[rng.integers(0, 4096, 900)].copy()
    # near duplicates: flip 1-3 bits
    bits = synth.unpack_bits(q)
    for i in range(0, 900, 3):
        for b in rng.integers(0, 352, rng.integers(1, 4)):
            bits[i, b] ^= True
    q = synth.pack_bits(bits)
    _check(mctx, oracle, q, t)


def test_self_match_initial_frame(mctx, oracle, synth):
    # src/lidar_odometry.cpp:187-194: the first frame is matched against itself
    d = synth.random_descriptors(600, seed=9, density=33)
    _check(mctx, oracle, d, d)


def test_all_ones_invalid_descriptors(mctx, oracle, synth):
    # NaN SHOT binarises to all 352 bits set (include/bshot_bits.h:166-260)
    q = synth.random_descriptors(300, seed=21)
    t = synth.random_descriptors(700, seed=22)
    ones = np.full(6, 0xFFFFFFFFFFFFFFFF, np.uint64)
    ones[5] = 0xFFFFFFFF
    q[5] = ones
    t[17] = ones
    t[400] = ones
    _check(mctx, oracle, q, t)


def test_extreme_popcounts_and_constant_targets(mctx, oracle, synth):
    """the corners of the key range: empty and full descriptors on both sides (distances 0 and 352), and a target set of
    identical records -- every distance ties, so the winner and the runner-up must be targets 0 and 1 whatever tile or
    split they fall into (first minimum of minVect, include/bshot_bits.h:6-20)"""
    ones = np.full(6, 0xFFFFFFFFFFFFFFFF, np.uint64)
    ones[5] = 0xFFFFFFFF
    q = synth.random_descriptors(300, seed=31)
    q[0] = 0
    q[1] = ones
    t = synth.random_descriptors(1000, seed=32)
    t[999] = 0
    t[128] = ones
    t[127] = ones
    _check(mctx, oracle, q, t)
    same = np.repeat(synth.random_descriptors(1, seed=33), 5000, axis=0)
    g = mctx.match(q, same)
    assert (g["left_idx"] == 0).all() and (g["left_idx2"] == 1).all()
    assert np.array_equal(g["left_dist"], g["left_dist2"])
    _check(mctx, oracle, q[:64], same[:700])


def test_empty_inputs(mctx, synth):
    q = synth.random_descriptors(10, seed=1)
    e = np.zeros((0, 6), np.uint64)
    g = mctx.match(q, e)
    assert (g["left_idx"] == -1).all() and (g["left_dist"] == -1).all()
    g = mctx.match(e, q)
    assert (g["right_idx"] == -1).all()
    pairs, dist = mctx.match_mutual(q, e)
    assert pairs.shape[0] == 0


def test_single_target_has_no_runner_up(mctx, synth):
    q = synth.random_descriptors(33, seed=1)
    t = synth.random_descriptors(1, seed=2)
    g = mctx.match(q, t)
    assert (g["left_idx"] == 0).all() and (g["left_idx2"] == -1).all() and (g["left_dist2"] == -1).all()


def test_sharded_merge_equals_single(mctx, bshot, oracle, synth):
    """8 emulated shards on one GPU: per-shard candidates + merge == single-pass result."""
    import torch
    nq, nt, shards = 1000, 40000, 8
    q = synth.random_descriptors(nq, seed=31)
    t = synth.random_descriptors(nt, seed=32)
    t[12345] = t[222]     # cross-shard tie: lowest global index must win
    t[39999] = q[5]
    t[77] = q[5]
    dq = torch.from_numpy(q.view(np.int64)).cuda()
    dt = torch.from_numpy(t.view(np.int64)).cuda()
    cands = torch.empty((shards, nq, 3), dtype=torch.int64, device="cuda")
    merged = torch.empty((nq, 3), dtype=torch.int64, device="cuda")
    per = (nt + shards - 1) // shards
    torch.cuda.synchronize()
    for r in range(shards):
        lo, hi = r * per, min(nt, (r + 1) * per)
        mctx.match_dev(dq.data_ptr(), nq, dt[lo:hi].data_ptr(), hi - lo, lo, True, cands[r].data_ptr())
    mctx.merge_cands_dev(cands.data_ptr(), shards, nq, merged.data_ptr())
    mctx.sync()
    rec = merged.cpu().numpy().view(bshot.CAND_DTYPE).reshape(nq)
    u = bshot.unpack_cands(rec)
    o = oracle.match(q, t, want_right=True)
    assert np.array_equal(u["idx1"], o["left_idx"])
    assert np.array_equal(u["dist1"], o["left_dist"])
    assert np.array_equal(u["idx2"], o["left_idx2"])
    assert np.array_equal(u["dist2"], o["left_dist2"])
    mutual_gpu = np.nonzero(u["rq"] == np.arange(nq))[0]
    opairs = oracle.mutual(o["left_idx"], o["right_idx"])
    assert np.array_equal(mutual_gpu, opairs[:, 0])


@pytest.mark.parametrize("kind", [0, 2], ids=["popc", "tensor-core-pipelined"])
def test_sharded_reverse_owned_equals_single(bshot, oracle, synth, kind):
    """4 emulated ranks (4 contexts on one GPU): shard search without rq, merge, reverse pass for the owned
    winners only, MAX-combine -> records identical to the single-pass oracle result."""
    import torch
    nq, nt, ranks = 700, 20000, 4
    q = synth.random_descriptors(nq, seed=41, density=40)
    t = synth.random_descriptors(nt, seed=42, density=40)
    t[19999] = q[3]
    t[10] = q[3]
    dq = torch.from_numpy(q.view(np.int64)).cuda()
    per = (nt + ranks - 1) // ranks
    ctxs = [bshot.Context(0, 1024, 1024, per) for _ in range(ranks)]
    try:
        for c in ctxs:
            c.set_matcher(kind)
        cands = torch.empty((ranks, nq, 3), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        for r, c in enumerate(ctxs):
            c.map_append(t[r * per:(r + 1) * per])
            c.match_shard_dev(dq.data_ptr(), nq, r * per, False, cands[r].data_ptr())
            c.sync()
        rqs = []
        merged_all = []
        for r, c in enumerate(ctxs):
            merged = torch.empty((nq, 3), dtype=torch.int64, device="cuda")
            rq = torch.empty(nq, dtype=torch.int32, device="cuda")
            c.merge_cands_dev(cands.data_ptr(), ranks, nq, merged.data_ptr())
            c.reverse_owned_dev(dq.data_ptr(), nq, r * per, merged.data_ptr(), rq.data_ptr())
            c.sync()
            rqs.append(rq)
            merged_all.append(merged)
        rq = torch.stack(rqs).max(dim=0).values            # the all-reduce(MAX)
        ctxs[0].apply_rq_dev(merged_all[0].data_ptr(), rq.data_ptr(), nq)
        ctxs[0].sync()
        rec = merged_all[0].cpu().numpy().view(bshot.CAND_DTYPE).reshape(nq)
    finally:
        for c in ctxs:
            c.close()
    u = bshot.unpack_cands(rec)
    o = oracle.match(q, t, want_right=True)
    assert np.array_equal(u["idx1"], o["left_idx"]) and np.array_equal(u["dist1"], o["left_dist"])
    assert np.array_equal(u["idx2"], o["left_idx2"])
    assert np.array_equal(u["rq"], o["right_idx"][o["left_idx"]])
    assert u["idx1"][3] == 10


@pytest.mark.parametrize("kind", [0, 2], ids=["popc", "tensor-core-pipelined"])
def test_sharded_peer_push_equals_single(bshot, oracle, synth, kind):
    """the peer-memory exchange on 4 emulated ranks (4 contexts on one GPU, every 'peer' buffer is a local
    buffer): every rank pushes its records into slot r of every rank's gather buffer, merges its own copy, runs the
    reverse pass for the winners it owns and pushes rq into every rank's array -> every rank ends with records
    identical to the single-pass oracle result."""
    import torch
    nq, nt, ranks = 700, 20000, 4
    q = synth.random_descriptors(nq, seed=51, density=40)
    t = synth.random_descriptors(nt, seed=52, density=40)
    t[19999] = q[5]
    t[12] = q[5]
    dq = torch.from_numpy(q.view(np.int64)).cuda()
    per = (nt + ranks - 1) // ranks
    ctxs = [bshot.Context(0, 1024, 1024, per) for _ in range(ranks)]
    try:
        for c in ctxs:
            c.set_matcher(kind)
        gather = [torch.zeros((ranks, nq, 3), dtype=torch.int64, device="cuda") for _ in range(ranks)]
        rqbuf = [torch.full((nq,), -7, dtype=torch.int32, device="cuda") for _ in range(ranks)]
        peer_g = torch.tensor([g.data_ptr() for g in gather], dtype=torch.int64, device="cuda")
        peer_r = torch.tensor([r.data_ptr() for r in rqbuf], dtype=torch.int64, device="cuda")
        local = torch.empty((nq, 3), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        for r, c in enumerate(ctxs):
            c.map_append(t[r * per:(r + 1) * per])
            c.match_shard_dev(dq.data_ptr(), nq, r * per, False, local.data_ptr())
            c.push_cands_dev(local.data_ptr(), nq, peer_g.data_ptr(), ranks, r)
            c.sync()
        for r in range(1, ranks):
            assert torch.equal(gather[r], gather[0])            # the barrier point: every rank holds all records
        merged = [torch.empty((nq, 3), dtype=torch.int64, device="cuda") for _ in range(ranks)]
        for r, c in enumerate(ctxs):
            c.merge_cands_dev(gather[r].data_ptr(), ranks, nq, merged[r].data_ptr())
            c.reverse_owned_push_dev(dq.data_ptr(), nq, r * per, merged[r].data_ptr(), peer_r.data_ptr(), ranks, r)
            c.sync()
        recs = []
        for r, c in enumerate(ctxs):                             # second barrier point, then every rank completes its records
            c.apply_rq_dev(merged[r].data_ptr(), rqbuf[r].data_ptr(), nq)
            c.sync()
            recs.append(merged[r].cpu().numpy().view(bshot.CAND_DTYPE).reshape(nq))
    finally:
        for c in ctxs:
            c.close()
    o = oracle.match(q, t, want_right=True)
    for rec in recs:
        u = bshot.unpack_cands(rec)
        assert np.array_equal(u["idx1"], o["left_idx"]) and np.array_equal(u["dist1"], o["left_dist"])
        assert np.array_equal(u["idx2"], o["left_idx2"])
        assert np.array_equal(u["rq"], o["right_idx"][o["left_idx"]])
    assert bshot.unpack_cands(recs[0])["idx1"][5] == 12


@pytest.mark.parametrize("kind", [0, 2], ids=["popc", "tensor-core-pipelined"])
def test_full_size_properties(gpu_ctx, bshot, synth, kind):
    """C4-sized shard (Q=10000 x T=1M): size-independent properties instead of the O(QT) oracle."""
    gpu_ctx.set_matcher(kind)
    try:
        _full_size_properties(gpu_ctx, bshot, synth)
    finally:
        gpu_ctx.set_matcher(-1)


def _full_size_properties(gpu_ctx, bshot, synth):
    nq, nt = 10000, 1 << 20
    t = synth.random_descriptors(nt, seed=7)
    rng = np.random.default_rng(5)
    src = rng.integers(0, nt, nq)
    q = t[src].copy()
    bits = synth.unpack_bits(q[: nq // 2])
    flip = rng.integers(0, 352, nq // 2)
    bits[np.arange(nq // 2), flip] ^= True          # first half: one bit flipped
    q[: nq // 2] = synth.pack_bits(bits)
    gpu_ctx.map_reset()
    gpu_ctx.map_append(t)
    u = bshot.unpack_cands(gpu_ctx.match_map(q, 0))
    # random 352-bit words: the planted source is the unique nearest neighbour
    assert np.array_equal(u["idx1"], src)
    assert (u["dist1"][: nq // 2] == 1).all() and (u["dist1"][nq // 2:] == 0).all()
    assert (u["dist2"] >= u["dist1"]).all() and (u["dist2"] > 100).all()
    # reverse check: each planted target's best query is the lowest query index that copied it
    first = {}
    for i, s in enumerate(src):
        first.setdefault(int(s), i)
    exp_rq = np.array([first[int(s)] for s in src])
    d_exp = np.where(exp_rq < nq // 2, 1, 0)
    # a later exact copy (dist 0) beats an earlier 1-bit-flipped copy
    for i, s in enumerate(src):
        if d_exp[i] == 1:
            later = [j for j in np.nonzero(src == s)[0] if j >= nq // 2]
            if later:
                exp_rq[i] = later[0]
    assert np.array_equal(u["rq"], exp_rq)

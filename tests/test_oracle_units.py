"""Pins the CPU oracle (oracle/bshot_oracle.cpp).  The reference has NO golden vectors (SURVEY 4,
8c: parity unpinned), so the oracle is pinned by (i) truth tables read off the reference source,
(ii) analytic cases, (iii) numpy / scipy cross-checks and (iv) an independent numpy restatement of
the PCL algorithms (tests/shot_numpy.py)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from shot_numpy import lrf_numpy, normal_numpy, seg_ratio_cv_numpy, shot_numpy


# ---- include/bshot_bits.h:144-278 ---------------------------------------------------------------
TRUTH = [
    ([0, 0, 0, 0], 0x0), ([1, 0, 0, 0], 0x1), ([0, 1, 0, 0], 0x2), ([0, 0, 1, 0], 0x4), ([0, 0, 0, 1], 0x8),
    ([.5, .5, 0, 0], 0x3), ([0, .5, .5, 0], 0x6), ([0, 0, .5, .5], 0xC), ([.5, 0, 0, .5], 0x9),
    ([0, .5, 0, .5], 0xA), ([.5, 0, .5, 0], 0x5), ([.34, .33, .33, 0], 0x7), ([0, .33, .34, .33], 0xE),
    ([.33, 0, .33, .34], 0xD), ([.33, .34, 0, .33], 0xB), ([.25, .25, .25, .25], 0xF),
    ([np.nan, 0, 0, 0], 0xF), ([np.nan] * 4, 0xF), ([0.91, 0.03, 0.03, 0.03], 0x1),
    ([0.45, 0.46, 0.05, 0.04], 0x3), ([1e-30, 0, 0, 0], 0x1),
]


def _nibble(bits6, j):
    return (int(bits6[(4 * j) // 64]) >> ((4 * j) % 64)) & 0xF


def test_bshot_truth_table(oracle):
    for j in (0, 15, 16, 87):            # nibbles that sit at word starts / ends
        for vec, exp in TRUTH:
            shot = np.zeros((1, 352), np.float32)
            shot[0, 4 * j: 4 * j + 4] = vec
            bits = oracle.bshot(shot)[0]
            assert _nibble(bits, j) == exp, (j, vec)
            other = [_nibble(bits, k) for k in range(88) if k != j]
            assert not any(other)


def test_bshot_nan_descriptor_is_all_ones(oracle):
    bits = oracle.bshot(np.full((1, 352), np.nan, np.float32))[0]
    assert [int(b) for b in bits[:5]] == [0xFFFFFFFFFFFFFFFF] * 5 and int(bits[5]) == 0xFFFFFFFF


def test_bshot_threshold_is_double_compare(oracle):
    # vec0 == float(0.9f * sum) would pass a float compare but not `vec[0] > 0.9 * sum` in double
    s = np.float32(1.0)
    v0 = np.float32(0.9)                     # float(0.9) > 0.9 (double) * 1.0 ? 0.89999998 < 0.9 -> no
    vec = [v0, s - v0, 0, 0]
    shot = np.zeros((1, 352), np.float32)
    shot[0, :4] = vec
    total = np.float32(np.float32(np.float32(vec[0]) + np.float32(vec[1])) + np.float32(0)) + np.float32(0)
    exp = 0x1 if float(v0) > 0.9 * float(total) else 0x3
    assert _nibble(oracle.bshot(shot)[0], 0) == exp == 0x3


def test_bitset352_layout_matches_libstdcxx(synth):
    """bshot_descriptor = std::bitset<352>: probe the real libstdc++ layout with g++ (Appendix B)."""
    src = r"""
#include <bitset>
#include <cstdio>
#include <cstring>
#include <cstdint>
int main(){ std::bitset<352> b; int s[] = {0,1,63,64,65,200,351}; for(int i: s) b.set(i);
 uint64_t w[6]; static_assert(sizeof(b)==48, "size"); memcpy(w,&b,48);
 for(int i=0;i<6;i++) printf("%llu\n",(unsigned long long)w[i]); }
"""
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "b.cpp")
        open(p, "w").write(src)
        subprocess.check_call(["/usr/bin/g++", "-O1", p, "-o", os.path.join(d, "b")])
        out = subprocess.check_output([os.path.join(d, "b")]).decode().split()
    bits = np.zeros((1, 352), bool)
    bits[0, [0, 1, 63, 64, 65, 200, 351]] = True
    assert [int(x) for x in out] == [int(x) for x in synth.pack_bits(bits)[0]]
    assert np.array_equal(synth.unpack_bits(synth.pack_bits(bits)), bits)


# ---- matching: src/lidar_odometry.cpp:212-242, include/bshot_bits.h:6-20 ------------------------
def _popcount_matrix(q, t):
    x = q[:, None, :] ^ t[None, :, :]
    return np.unpackbits(x.view(np.uint8), axis=-1).sum(-1).astype(np.int32)


def test_match_first_minimum_and_mutual(oracle, synth):
    q = synth.random_descriptors(40, seed=1, density=30)
    t = synth.random_descriptors(70, seed=2, density=30)
    t[50] = t[3]                            # duplicate target: index 3 must win (strict '<')
    q[7] = t[3]
    d = _popcount_matrix(q, t)
    m = oracle.match(q, t)
    assert np.array_equal(m["left_idx"], d.argmin(1))       # np.argmin = first minimum
    assert np.array_equal(m["left_dist"], d.min(1))
    assert np.array_equal(m["right_idx"], d.argmin(0))
    assert m["left_idx"][7] == 3 and m["left_dist"][7] == 0
    # runner-up: second in (distance, index) order
    order = np.lexsort((np.broadcast_to(np.arange(70), d.shape), d), axis=1)
    assert np.array_equal(m["left_idx2"], order[:, 1])
    pairs = oracle.mutual(m["left_idx"], m["right_idx"])
    exp = [(i, j) for i, j in enumerate(d.argmin(1)) if d.argmin(0)[j] == i]
    assert [tuple(p) for p in pairs] == exp


def test_initial_frame_self_match(oracle, synth):
    # src/lidar_odometry.cpp:187-194: every keypoint is its own mutual NN unless an earlier duplicate exists
    d = synth.random_descriptors(64, seed=3)
    d[40] = d[10]
    m = oracle.match(d, d)
    exp = np.arange(64)
    exp[40] = 10
    assert np.array_equal(m["left_idx"], exp) and (m["left_dist"] == 0).all()
    pairs = oracle.mutual(m["left_idx"], m["right_idx"])
    assert 40 not in pairs[:, 0] and len(pairs) == 63


# ---- radius search: Appendix A.1 ---------------------------------------------------------------
def _brute_radius(pts, q, r, max_nn=0):
    d = (q - pts).astype(np.float32)
    sq = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    idx = np.nonzero(sq < np.float32(r * r))[0]
    order = np.lexsort((idx, sq[idx]))
    idx = idx[order]
    if max_nn:
        idx = idx[:max_nn]
    return idx.astype(np.int32), sq[idx]


def test_radius_search_exact_sorted_capped(oracle):
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(0)
    pts = rng.uniform(-5000, 5000, (5000, 3)).astype(np.float32)
    pts[100] = pts[7]                       # exact duplicate -> tie ordered by index
    oc = oracle.Cloud(pts)
    tree = cKDTree(pts.astype(np.float64))
    for qi in (0, 7, 100, 2500):
        for r, cap in ((800.0, 0), (2000.0, 0), (2000.0, 50), (3000.0, 300)):
            idx, sqd = oc.radius_search(pts[qi], r, cap)
            bi, bs = _brute_radius(pts, pts[qi], r, cap)
            assert np.array_equal(idx, bi) and np.array_equal(sqd, bs)
            if cap == 0:
                ball = tree.query_ball_point(pts[qi].astype(np.float64), r * (1 - 1e-6))
                assert set(ball) <= set(idx.tolist())
    # query far outside the cloud / non-finite query
    assert len(oc.radius_search(np.array([1e6, 0, 0], np.float32), 1000.0)[0]) == 0
    assert len(oc.radius_search(np.array([np.nan, 0, 0], np.float32), 1000.0)[0]) == 0


# ---- eigen solvers ------------------------------------------------------------------------------
def test_eigh3_vs_numpy(oracle):
    rng = np.random.default_rng(1)
    for _ in range(200):
        a = rng.normal(size=(3, 3)) * 10 ** rng.uniform(-3, 6)
        m = a @ a.T
        w, v = oracle.eigh3(m)
        wn, vn = np.linalg.eigh(m)
        assert np.allclose(w, wn, rtol=1e-12, atol=1e-12 * abs(wn).max())
        for c in range(3):
            assert abs(abs(v[:, c] @ vn[:, c]) - 1) < 1e-9
    w, v = oracle.eigh3(np.diag([3.0, 1.0, 2.0]))
    assert np.allclose(w, [1, 2, 3])


def test_eigen33_smallest_vs_numpy(oracle):
    rng = np.random.default_rng(2)
    for _ in range(100):
        a = rng.normal(size=(3, 3))
        m = (a @ np.diag([1.0, 0.5, 0.05]) @ a.T).astype(np.float32)
        m = ((m + m.T) / 2).astype(np.float32)
        ev, vec = oracle.eigen33_smallest(m)
        wn, vn = np.linalg.eigh(m.astype(np.float64))
        assert abs(ev - wn[0]) < 1e-4 * wn[2]
        assert abs(abs(vec @ vn[:, 0]) - 1) < 1e-3


# ---- independent numpy restatement (tests/shot_numpy.py) -----------------------------------------
@pytest.fixture(scope="module")
def small_cloud():
    rng = np.random.default_rng(4)
    a = rng.uniform(-1500, 1500, (400, 3))
    a[:, 2] = 0.15 * a[:, 0] + rng.normal(0, 30, 400)          # noisy tilted plane
    b = rng.uniform(-1500, 1500, (300, 3))
    b[:, 0] = 400 + rng.normal(0, 20, 300)                      # a wall
    return np.concatenate([a, b]).astype(np.float32)


def test_seg_ratio_vs_numpy(oracle, small_cloud):
    oc = oracle.Cloud(small_cloud)
    r = oc.seg_ratio(600.0, 40, oracle.SR_CV, threads=2)
    for i in range(0, len(small_cloud), 37):
        assert r[i] == seg_ratio_cv_numpy(small_cloud, i, 600.0, 40) or (
            np.isnan(r[i]) and np.isnan(seg_ratio_cv_numpy(small_cloud, i, 600.0, 40)))


def test_seg_ratio_skips_origin_and_select(oracle):
    pts = np.array([[0, 0, 0], [100, 0, 0], [200, 10, 0], [300, 0, 5], [150, 50, 0]], np.float32)
    r = oracle.Cloud(pts).seg_ratio(1000.0, 300, oracle.SR_CV)
    assert np.isnan(r[0]) and not np.isnan(r[1:]).any()
    ratio = np.array([0.5, np.nan, 0.9, 0.5, 0.7, 0.9], np.float32)
    idx, rat = oracle.select_keypoints(ratio, 3, oracle.TIE_DETERMINISTIC)
    assert list(idx) == [4, 5, 2] and list(rat) == [np.float32(0.7), np.float32(0.9), np.float32(0.9)]
    idx, _ = oracle.select_keypoints(ratio, 10, oracle.TIE_DETERMINISTIC)    # fewer than K valid: all
    assert sorted(idx) == [0, 2, 3, 4, 5]
    idx_s, rat_s = oracle.select_keypoints(ratio, 3, oracle.TIE_STDSORT)
    assert sorted(rat_s) == sorted(rat)


def test_normals_vs_numpy_and_plane(oracle, small_cloud):
    oc = oracle.Cloud(small_cloud)
    q = small_cloud[:50]
    n = oc.normals(q, 500.0, 30)
    for i in range(0, 50, 7):
        ref = normal_numpy(small_cloud, q[i], 500.0, 30)
        assert np.allclose(n[i], ref, atol=2e-3), (i, n[i], ref)
    # points exactly on z = 0: normal is +-z and flipped towards the origin viewpoint
    g = np.stack(np.meshgrid(np.arange(-10, 11), np.arange(-10, 11)), -1).reshape(-1, 2) * 50.0
    plane = np.concatenate([g, np.full((len(g), 1), -1000.0)], 1).astype(np.float32)
    n = oracle.Cloud(plane).normals(plane[200:210], 300.0, 300)
    assert np.allclose(np.abs(n[:, 2]), 1.0, atol=1e-3) and (n[:, 2] > 0).all()
    assert np.allclose(n[:, 3], 0.0, atol=1e-3)
    # fewer than 3 neighbours -> NaN ; no neighbour -> NaN (include/bshot_bits.h:67-74)
    lone = np.array([[0, 0, 1], [5000, 0, 0], [5001, 0, 0]], np.float32)
    n = oracle.Cloud(lone).normals(lone, 10.0, 300)
    assert np.isnan(n).all()


def test_lrf_vs_numpy(oracle, small_cloud):
    oc = oracle.Cloud(small_cloud)
    kp = small_cloud[::50]
    rf, valid = oc.lrf(kp, 700.0)
    for i in range(len(kp)):
        ref, nv = lrf_numpy(small_cloud, kp[i], 700.0)
        assert nv == valid[i]
        assert np.allclose(rf[i], ref, atol=1e-5, equal_nan=True), i
    ok = ~np.isnan(rf).any(1)
    x, y, z = rf[ok, :3], rf[ok, 3:6], rf[ok, 6:]
    assert np.allclose(np.cross(z, x), y, atol=1e-6)
    assert np.allclose(np.linalg.norm(x, axis=1), 1, atol=1e-6)


def test_lrf_too_few_neighbours_is_nan(oracle):
    pts = np.array([[0, 0, 0], [10, 0, 0], [0, 10, 0], [0, 0, 10], [5, 5, 5]], np.float32)
    rf, valid = oracle.Cloud(pts).lrf(pts[:1], 100.0)          # 4 valid neighbours (< 5)
    assert valid[0] == 4 and np.isnan(rf).all()
    pts = np.concatenate([pts, [[7, 1, 2]]]).astype(np.float32)
    rf, valid = oracle.Cloud(pts).lrf(pts[:1], 100.0)
    assert valid[0] == 5 and not np.isnan(rf).any()


def test_shot_vs_numpy_and_reference_quirk(oracle, synth, small_cloud):
    oc = oracle.Cloud(small_cloud)
    kp = small_cloud[::70]
    normals = oc.normals(small_cloud, 400.0, 40)
    shot, rf, nn, total = oc.shot(kp, normals, 700.0)
    assert total == nn.sum()
    for i in range(len(kp)):
        ref = shot_numpy(small_cloud, kp[i], 700.0, normals, rf[i])
        assert np.allclose(shot[i], ref, atol=2e-6, equal_nan=True), i
    assert np.allclose(np.linalg.norm(shot[~np.isnan(shot).any(1)], axis=1), 1.0, atol=1e-5)
    # REFERENCE quirk: zero normals => all mass in slot 5 of each volume
    d = oc.compute_descriptors(kp, 700.0, 40, oracle.MODE_REFERENCE, want_normals=True)
    assert (d["normals"][len(kp):] == 0).all()
    s = d["shot"][~np.isnan(d["shot"]).any(1)].reshape(-1, 32, 11)
    zero_normal_mass = s[:, :, 5].sum() / s.sum()
    assert zero_normal_mass > 0.9
    bits = synth.unpack_bits(d["bits"][~np.isnan(d["shot"]).any(1)])
    assert bits.sum(1).max() <= 48


def test_shot_rigid_motion_invariance(oracle, small_cloud):
    """rotating + translating cloud, keypoints and normals rotates the LRF and leaves SHOT unchanged"""
    th = 0.7
    Rm = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    Rx = np.array([[1, 0, 0], [0, np.cos(0.3), -np.sin(0.3)], [0, np.sin(0.3), np.cos(0.3)]])
    Rm = Rm @ Rx
    moved = (small_cloud.astype(np.float64) @ Rm.T + [120.0, -50.0, 30.0]).astype(np.float32)
    oc0, oc1 = oracle.Cloud(small_cloud), oracle.Cloud(moved)
    sel = np.arange(0, len(small_cloud), 90)
    n0 = oc0.normals(small_cloud, 400.0, 0)
    # orient consistently by transporting the normals instead of re-flipping towards the new origin
    n1 = n0.copy()
    n1[:, :3] = (n0[:, :3].astype(np.float64) @ Rm.T).astype(np.float32)
    s0, rf0, _, _ = oc0.shot(small_cloud[sel], n0, 700.0)
    s1, rf1, _, _ = oc1.shot(moved[sel], n1, 700.0)
    ok = ~np.isnan(s0).any(1)
    assert np.allclose(rf0[ok].reshape(-1, 3, 3) @ Rm.T, rf1[ok].reshape(-1, 3, 3), atol=2e-3)
    assert np.abs(s0[ok] - s1[ok]).max() < 5e-3


def test_planted_duplicate_descriptors_have_distance_zero(oracle, small_cloud):
    """two copies of a patch, translated by an amount that is exact in fp32 (coordinates quantised
    to 1/4 mm), with transported normals -> bit-identical descriptors -> Hamming distance 0"""
    base = (np.round(small_cloud * 4) / 4).astype(np.float32)
    far = base + np.array([16384, 0, 0], np.float32)
    assert np.array_equal((far - np.array([16384, 0, 0], np.float32)), base)
    both = np.concatenate([base, far])
    oc = oracle.Cloud(both)
    n = oracle.Cloud(base).normals(base, 400.0, 40)
    normals = np.concatenate([n, n])
    sel = np.arange(0, 700, 100)
    kp = np.concatenate([base[sel], far[sel]])
    shot, rf, nn, _ = oc.shot(kp, normals, 600.0)
    assert np.array_equal(shot[:7], shot[7:], equal_nan=True)
    bits = oracle.bshot(shot)
    m = oracle.match(bits[:7], bits[7:])
    assert (m["left_dist"] == 0).all()


def _bin_centre_cloud(R, placements):
    """keypoint at the origin (identity LRF given explicitly) + one neighbour per (sector, upper, outer) placed at
    the CENTRE of its spatial volume: azimuth -7pi/8 + sel*pi/4, inclination pi/4 (upper) or 3pi/4 (lower) from +z,
    distance R/4 (inner husk) or 3R/4 (outer husk)."""
    pts = [(0.0, 0.0, 0.0)]
    for sel, upper, outer in placements:
        phi = -7.0 * np.pi / 8.0 + sel * np.pi / 4.0
        theta = np.pi / 4.0 if upper else 3.0 * np.pi / 4.0
        d = 0.75 * R if outer else 0.25 * R
        pts.append((d * np.sin(theta) * np.cos(phi), d * np.sin(theta) * np.sin(phi), d * np.cos(theta)))
    return np.asarray(pts, np.float32)


def test_shot_known_answer_at_bin_centres(oracle):
    """Hand-computed SHOT352: a neighbour that sits at the centre of its volume in all four
    interpolated dimensions (cosine bin, radial husk, inclination, azimuth sector) puts its whole weight
    1 + 1 + 1 + 1 = 4 into ONE bin, index (sel*4 + outer*2 + upper)*11 + step with step = 10 for a normal parallel
    to the z axis of the frame and step = 5 for a zero normal (the reference quirk).  N such neighbours in N
    different volumes give a normalised histogram with exactly N entries 1/sqrt(N), and B-SHOT sets exactly those
    bits (a lone non-zero value in its group of four exceeds 0.9 * sum)."""
    R = 3000.0
    placements = [(0, 1, 0), (1, 1, 0), (2, 0, 0), (3, 0, 1), (4, 1, 1), (5, 0, 0), (6, 1, 1), (7, 0, 1)]
    pts = _bin_centre_cloud(R, placements)
    oc = oracle.Cloud(pts)
    rf_identity = np.array([[1, 0, 0, 0, 1, 0, 0, 0, 1]], np.float32)
    for normal, step in (((0.0, 0.0, 1.0), 10), ((0.0, 0.0, 0.0), 5), ((0.0, 0.0, -1.0), 0)):
        normals = np.zeros((len(pts), 4), np.float32)
        normals[:, :3] = normal
        shot, rf, nn, total = oc.shot(pts[:1], normals, R, rf_in=rf_identity)
        assert nn[0] == len(pts) and total == len(pts)
        expect = np.zeros(352, np.float32)
        for sel, upper, outer in placements:
            expect[(sel * 4 + outer * 2 + upper) * 11 + step] = 1.0 / np.sqrt(len(placements))
        # float32 positions are not exactly on the centres: a leak of ~1e-7 into the neighbouring bins is expected
        assert np.abs(shot[0] - expect).max() < 2e-6, np.abs(shot[0] - expect).max()
        bits = oracle.bshot(shot)
        from_bits = np.unpackbits(bits.view(np.uint8), bitorder="little")[:352].astype(bool)
        assert from_bits[expect > 0.1].all()
        # B-SHOT is scale free per group of four: a 1e-8 leak that is alone in its group sets its bit as well
        # (include/bshot_bits.h:171-178), so extra bits may only sit on non-zero leak bins
        extra = from_bits & ~(expect > 0.1)
        assert (shot[0][extra] != 0).all() and (np.abs(shot[0][extra]) < 2e-6).all()


def test_shot_known_answer_between_two_sectors(oracle):
    """A neighbour exactly on the border of two azimuth sectors (angular offset 0.5) splits its azimuth share half
    and half: own bin 1 (cosine) + 1 (radial) + 1 (inclination) + 0.5, the sector on the other side of the border
    0.5; with a second, mirrored neighbour the two 3.5 / 0.5 pairs normalise to 3.5/5 and 0.5/5."""
    R = 3000.0
    pts = [(0.0, 0.0, 0.0)]
    for sign in (+1.0, -1.0):                      # borders between sectors 3|4 (phi = 0) and 7|0 (phi = pi)
        phi = 0.0 if sign > 0 else np.pi
        theta, d = np.pi / 4.0, 0.25 * R
        pts.append((d * np.sin(theta) * np.cos(phi), 0.0, d * np.cos(theta)))
    for k in range(3):                             # filler so that the neighbourhood has >= 5 points; placed at centres
        pts.append(tuple(_bin_centre_cloud(R, [(2 * k + 1, 0, 1)])[1]))
    pts = np.asarray(pts, np.float32)
    oc = oracle.Cloud(pts)
    normals = np.zeros((len(pts), 4), np.float32)
    normals[:, 2] = 1.0
    shot, _, nn, _ = oc.shot(pts[:1], normals, R, rf_in=np.array([[1, 0, 0, 0, 1, 0, 0, 0, 1]], np.float32))
    h = shot[0]
    raw = h / h.max() * 4.0                        # undo the normalisation: the filler bins hold exactly 4
    nz = {int(i): float(raw[i]) for i in np.nonzero(np.abs(raw) > 1e-4)[0]}
    fillers = {(s * 4 + 2 + 0) * 11 + 10 for s in (1, 3, 5)}
    assert fillers <= set(nz) and all(abs(nz[i] - 4.0) < 1e-4 for i in fillers)
    split = {i: v for i, v in nz.items() if i not in fillers}
    # y == 0: PCL's sector rule puts x > 0 into sector 3 or 4 and x < 0 into 7 or 0; either way the pair of bins that
    # share the border hold 3.5 and 0.5
    assert len(split) == 4
    vals = sorted(split.values())
    assert np.allclose(vals, [0.5, 0.5, 3.5, 3.5], atol=1e-4)
    for (a, b) in ((3, 4), (7, 0)):
        ia, ib = (a * 4 + 1) * 11 + 10, (b * 4 + 1) * 11 + 10
        assert {ia, ib} <= set(split) and abs(split[ia] + split[ib] - 4.0) < 1e-4


def test_seg_ratio_known_answer_on_a_line(oracle):
    """Hand-computed CV seg-ratio (src/lidar_odometry.cpp:76-97) for five collinear points 100 mm apart: the end
    points see all neighbours on one side of the plane through themselves perpendicular to (p - centroid) ->
    1 - 0/4 = 1; the second point has one neighbour on one side and three on the other -> 1 - 1/3; the middle point
    coincides with the centroid, every dot product is 0 -> 0/0 = NaN (the reference skips it, :121)."""
    x = 1000.0 + 100.0 * np.arange(5, dtype=np.float32)
    pts = np.stack([x, np.full(5, 50.0, np.float32), np.full(5, -20.0, np.float32)], 1)
    r = oracle.Cloud(pts).seg_ratio(1000.0, 300, oracle.SR_CV)
    assert r[0] == 1.0 and r[4] == 1.0
    assert r[1] == np.float32(1.0) - np.float32(1.0) / np.float32(3.0) and r[3] == r[1]
    assert np.isnan(r[2])
    # capped neighbourhood: max_nn = 3 keeps the 3 nearest (self + the two at 100 mm, for an end point self + 100 + 200)
    r3 = oracle.Cloud(pts).seg_ratio(1000.0, 3, oracle.SR_CV)
    assert r3[0] == 1.0 and np.isnan(r3[2]) and r3[4] == 1.0


def test_lrf_known_answer_on_the_axes(oracle):
    """Hand-computed SHOT LRF: neighbours on the coordinate axes give a diagonal weighted
    covariance, so x^ = +-e_x (largest eigenvalue) and z^ = +-e_z (smallest) for the distances below.  The sign rule
    counts v.axis >= 0 as '+', so the points on the OTHER axes (projection 0) vote '+' as well: an axis flips only
    when the neighbours on its negative side outnumber its positive side plus all zero projections.  y^ = z^ x x^."""
    R = 3000.0

    def cloud(xs, ys, zs):
        pts = [(0.0, 0.0, 0.0)] + [(x, 0.0, 0.0) for x in xs] + [(0.0, y, 0.0) for y in ys] + [(0.0, 0.0, z) for z in zs]
        return np.asarray(pts, np.float32)

    neg8 = lambda d: [-d * (1.0 - 0.03 * k) for k in range(8)]
    cases = [
        # no flip: xx > yy > zz, majorities on the positive sides
        (cloud([2000, 1800, -2000], [1500, -1500], [1000, 900, -1000]), (1, 0, 0), (0, 0, 1)),
        # eight neighbours on -x against 1 (+x) + 4 zero projections: x^ flips, z^ does not
        (cloud([2000] + neg8(2000), [1500, -1500], [1000, -1000]), (-1, 0, 0), (0, 0, 1)),
        # eight neighbours on -z (close, so zz stays the smallest) against 1 (+z) + 4 zeros: z^ flips, x^ does not
        (cloud([2000, -2000], [1500, -1500], [500] + neg8(500)), (1, 0, 0), (0, 0, -1)),
    ]
    for pts, x_axis, z_axis in cases:
        rf, valid = oracle.Cloud(pts).lrf(pts[:1], R)
        assert valid[0] == len(pts) - 1                        # the keypoint itself is not a neighbour (A.4)
        y_axis = np.cross(np.asarray(z_axis, float), np.asarray(x_axis, float))
        assert np.allclose(rf[0], np.concatenate([x_axis, y_axis, z_axis]), atol=1e-6), (rf[0], x_axis, z_axis)
    # fewer than 5 neighbours: NaN frame (A.4)
    few = cases[0][0][:5]
    rf, valid = oracle.Cloud(few).lrf(few[:1], R)
    assert valid[0] == 4 and np.isnan(rf[0]).all()


def test_radius_search_boundary_is_strict(oracle):
    """FLANN's RadiusResultSet keeps dist < radius^2 (strict): a point exactly on the sphere is not a
    neighbour; equal distances come out in index order; max_nn keeps the nearest hits."""
    pts = np.array([(10, 0, 0), (3010, 0, 0), (3009, 0, 0), (10, 500, 0), (10, -500, 0), (10, 0, 500)], np.float32)
    oc = oracle.Cloud(pts)
    idx, sqd = oc.radius_search(pts[0], 3000.0, 0)
    assert list(idx) == [0, 3, 4, 5, 2]                       # index 1 sits exactly at 3000 mm: excluded
    assert list(sqd) == [0.0, 250000.0, 250000.0, 250000.0, 2999.0 ** 2]
    idx3, _ = oc.radius_search(pts[0], 3000.0, 3)
    assert list(idx3) == [0, 3, 4]                             # the 3 nearest; ties broken by index

"""Independent numpy restatement of the PCL algorithms behind the reference's front end
(SURVEY.md Appendix A), used ONLY to cross-check oracle/bshot_oracle.cpp on small clouds.
Written array-at-a-time from the published algorithm descriptions, deliberately not sharing code
with the C++ oracle.  Eigen decompositions use numpy.linalg.eigh."""
import numpy as np

f32 = np.float32


def radius_search(pts, q, r, max_nn=0):
    d = (np.asarray(q, f32) - pts).astype(f32)
    sq = ((d[:, 0] * d[:, 0]).astype(f32) + (d[:, 1] * d[:, 1]).astype(f32)).astype(f32)
    sq = (sq + (d[:, 2] * d[:, 2]).astype(f32)).astype(f32)
    idx = np.nonzero(sq < f32(float(r) * float(r)))[0]
    idx = idx[np.lexsort((idx, sq[idx]))]
    if max_nn:
        idx = idx[:max_nn]
    return idx, sq[idx]


def seg_ratio_cv_numpy(pts, i, r, max_nn):
    """src/lidar_odometry.cpp:61-97"""
    sp = pts[i]
    if (sp == 0).all():
        return np.nan
    idx, _ = radius_search(pts, sp, r, max_nn)
    acc = np.zeros(3, f32)
    for j in idx:                               # fp32 running sum in neighbour order
        acc = (acc + pts[j]).astype(f32)
    ct = (acc / f32(len(idx))).astype(f32)
    v = (sp - ct).astype(f32)
    rel = (pts[idx] - sp).astype(f32)
    dots = ((v[0] * rel[:, 0]).astype(f32) + (v[1] * rel[:, 1]).astype(f32)).astype(f32)
    dots = (dots + (v[2] * rel[:, 2]).astype(f32)).astype(f32)
    pos, neg = f32((dots > 0).sum()), f32((dots < 0).sum())
    if max(pos, neg) == 0:
        return np.nan
    return f32(1) - min(pos, neg) / max(pos, neg)


def normal_numpy(pts, q, r, max_nn):
    """pcl::computePointNormal + flipNormalTowardsViewpoint(0,0,0) -- mathematically (float64
    covariance), so it agrees with the fp32 oracle only to the conditioning of the input."""
    idx, _ = radius_search(pts, q, r, max_nn)
    if len(idx) < 3:
        return np.full(4, np.nan, f32)
    p = pts[idx].astype(np.float64)
    c = np.cov(p.T, bias=True)
    w, v = np.linalg.eigh(c)
    n = v[:, 0]
    if (-np.asarray(q, np.float64)) @ n < 0:
        n = -n
    curv = abs(w[0] / w.sum()) if w.sum() != 0 else 0.0
    return np.array([n[0], n[1], n[2], curv], f32)


def lrf_numpy(pts, q, r):
    """SHOTLocalReferenceFrameEstimation::getLocalRF"""
    q = np.asarray(q, f32)
    idx, sq = radius_search(pts, q, r, 0)
    keep = ~(pts[idx] == q).all(1)
    idx, sq = idx[keep], sq[keep]
    nv = len(idx)
    if nv < 5:
        return np.full(9, np.nan, f32), nv
    v = (pts[idx] - q).astype(f32).astype(np.float64)
    w = float(r) - np.sqrt(sq.astype(np.float64))
    m = (v * w[:, None]).T @ v / w.sum()
    ev, evec = np.linalg.eigh(m)
    x, z = evec[:, 2].copy(), evec[:, 0].copy()
    for ax in (x, z):
        s = 2 * int((v @ ax >= 0).sum()) - nv
        if s == 0:
            med = nv // 2
            cnt = int((v[med - 2: med + 3] @ ax > 0).sum())
            if cnt < 3:
                ax *= -1
        elif s < 0:
            ax *= -1
    xf, zf = x.astype(f32), z.astype(f32)
    yf = np.array([zf[1] * xf[2] - zf[2] * xf[1], zf[2] * xf[0] - zf[0] * xf[2],
                   zf[0] * xf[1] - zf[1] * xf[0]], f32)
    return np.concatenate([xf, yf, zf]).astype(f32), nv


def shot_numpy(pts, q, r, normals4, rf):
    """SHOTEstimation::computePointSHOT (shape only, 10 bins, 32 volumes) for one keypoint"""
    q = np.asarray(q, f32)
    idx, sq = radius_search(pts, q, r, 0)
    if len(idx) < 5 or np.isnan(rf).any():
        return np.full(352, np.nan, f32)
    xa, ya, za = rf[0:3].astype(f32), rf[3:6].astype(f32), rf[6:9].astype(f32)
    H = np.zeros(352, f32)

    def dot(a, b):
        return f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))

    def add(i, val):
        H[i] = f32(H[i] + f32(val))

    R12, R14, R34 = r / 2.0, r / 4.0, 3.0 * r / 4.0
    for s, sqd in zip(idx, sq):
        n = normals4[s, :3]
        if not np.isfinite(n).all():
            continue
        c = min(1.0, max(-1.0, float(dot(n, za))))
        b = (1.0 + c) * 10 / 2
        d = float(np.sqrt(np.float64(sqd)))
        if d < 1e-15:
            continue
        delta = (pts[s] - q).astype(f32)
        x, y, z = float(dot(delta, xa)), float(dot(delta, ya)), float(dot(delta, za))
        # 8 azimuth sectors x {lower, upper} x {inner, outer}
        bit4 = 1 if (y > 0 or (y == 0 and x < 0)) else 0
        bit3 = (1 - bit4) if (x > 0 or (x == 0 and y > 0)) else bit4
        v = ((bit4 << 3) + (bit3 << 2)) << 1
        if x * y > 0 or x == 0:
            v += 0 if abs(x) >= abs(y) else 4
        else:
            v += 4 if abs(x) > abs(y) else 0
        v += 1 if z > 0 else 0
        v += 2 if d > R12 else 0
        step = int(np.floor(b + 0.5))
        beta = b - step
        w = 1 - abs(beta)
        if beta > 0:
            add(v * 11 + (step + 1) % 10, beta)
        else:
            add(v * 11 + (step - 1 + 10) % 10, -beta)
        if d > R12:
            rho = (d - R34) / R12
            if d > R34:
                w += 1 - rho
            else:
                w += 1 + rho
                add((v - 2) * 11 + step, -rho)
        else:
            rho = (d - R14) / R12
            if d < R14:
                w += 1 + rho
            else:
                w += 1 - rho
                add((v + 2) * 11 + step, rho)
        th = float(np.arccos(min(1.0, max(-1.0, z / d))))
        if th > np.pi / 2 or (abs(th - np.pi / 2) < 1e-30 and z <= 0):
            io = (th - 3 * np.pi / 4) / (np.pi / 2)
            if th > 3 * np.pi / 4:
                w += 1 - io
            else:
                w += 1 + io
                add((v + 1) * 11 + step, -io)
        else:
            io = (th - np.pi / 4) / (np.pi / 2)
            if th < np.pi / 4:
                w += 1 + io
            else:
                w += 1 - io
                add((v - 1) * 11 + step, io)
        if x != 0 or y != 0:
            phi = float(np.arctan2(y, x))
            sel = v >> 2
            al = (phi - (-7 * np.pi / 8 + (np.pi / 4) * sel)) / (np.pi / 4)
            al = max(-0.5, min(al, 0.5))
            if al > 0:
                w += 1 - al
                add(((v + 4) % 32) * 11 + step, al)
            else:
                w += 1 + al
                add(((v - 4 + 32) % 32) * 11 + step, -al)
        add(v * 11 + step, w)
    norm = np.sqrt(np.sum((H * H).astype(f32).astype(np.float64)))
    return (H / f32(norm)).astype(f32)

"""Seeded synthetic Velodyne-shaped scans -- workload generator for tests/bench.

Units are millimetres, like the reference (src/preprocess.cpp:46).  The scene is a 200 m x 30 m
street canyon: two side walls, two end walls, 40 boxes, 30 vertical cylinders.  There is no ground
plane (the reference removes the ground before the path, src/preprocess.cpp:73-166), so downward
beams continue to the structure behind.  HDL-32E uses the beam table of the reference's capture
class (include/VelodyneCapture.h:572); HDL-64E is 64 beams uniform in [-24.8, +2.0] degrees.
"""
import numpy as np

SCENE_SEED = 20260118

HDL32E_LUT = np.array(
    [-30.67, -9.3299999, -29.33, -8.0, -28, -6.6700001, -26.67, -5.3299999, -25.33, -4.0, -24.0,
     -2.6700001, -22.67, -1.33, -21.33, 0.0, -20.0, 1.33, -18.67, 2.6700001, -17.33, 4.0, -16,
     5.3299999, -14.67, 6.6700001, -13.33, 8.0, -12.0, 9.3299999, -10.67, 10.67])

SENSORS = {
    # name: (vertical angles deg, azimuth steps, max range mm)
    "hdl32e": (HDL32E_LUT, 2170, 70000.0),
    "hdl64e": (np.linspace(-24.8, 2.0, 64), 1875, 120000.0),
}


def _scene(seed=SCENE_SEED):
    rng = np.random.default_rng(seed)
    ground = -1900.0
    boxes = []
    for _ in range(40):
        sx, sy, sz = rng.uniform(1000.0, 6000.0, 3)
        cx = rng.uniform(-95000.0, 95000.0)
        cy = rng.uniform(-13000.0, 13000.0)
        if abs(cx) < 4000 and abs(cy) < 4000:  # keep the sensor start clear
            cy = np.sign(cy + 1e-3) * 8000.0
        boxes.append((cx - sx / 2, cx + sx / 2, cy - sy / 2, cy + sy / 2, ground, ground + sz))
    cyls = []
    for _ in range(30):
        r = rng.uniform(150.0, 500.0)
        cx = rng.uniform(-95000.0, 95000.0)
        cy = rng.uniform(-14000.0, 14000.0)
        if abs(cx) < 3000 and abs(cy) < 3000:
            cx += 6000.0
        h = rng.uniform(3000.0, 10000.0)
        cyls.append((cx, cy, r, ground, ground + h))
    return np.array(boxes), np.array(cyls)


def _raycast(o, d, boxes, cyls):
    """nearest positive hit distance per ray (inf if none). o:(3,), d:(n,3) unit."""
    n = d.shape[0]
    t = np.full(n, np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        # side walls y = +-15 m, |x| <= 100 m ; end walls x = +-100 m, |y| <= 15 m
        for yw in (-15000.0, 15000.0):
            tt = (yw - o[1]) / d[:, 1]
            x = o[0] + tt * d[:, 0]
            ok = (tt > 0) & (np.abs(x) <= 100000.0)
            t = np.where(ok & (tt < t), tt, t)
        for xw in (-100000.0, 100000.0):
            tt = (xw - o[0]) / d[:, 0]
            y = o[1] + tt * d[:, 1]
            ok = (tt > 0) & (np.abs(y) <= 15000.0)
            t = np.where(ok & (tt < t), tt, t)
        inv = 1.0 / d
        for b in boxes:  # slab method
            t0 = (np.array([b[0], b[2], b[4]]) - o) * inv
            t1 = (np.array([b[1], b[3], b[5]]) - o) * inv
            tmin = np.minimum(t0, t1).max(axis=1)
            tmax = np.maximum(t0, t1).min(axis=1)
            ok = (tmax >= tmin) & (tmin > 0)
            t = np.where(ok & (tmin < t), tmin, t)
        for c in cyls:  # vertical finite cylinder, side surface only
            ox, oy = o[0] - c[0], o[1] - c[1]
            a = d[:, 0] ** 2 + d[:, 1] ** 2
            bq = 2 * (ox * d[:, 0] + oy * d[:, 1])
            cq = ox * ox + oy * oy - c[2] ** 2
            disc = bq * bq - 4 * a * cq
            tt = (-bq - np.sqrt(np.maximum(disc, 0))) / (2 * a)
            z = o[2] + tt * d[:, 2]
            ok = (disc > 0) & (tt > 0) & (z >= c[3]) & (z <= c[4])
            t = np.where(ok & (tt < t), tt, t)
    return t


def make_scan(sensor="hdl32e", frame=0, noise_mm=20.0, scene_seed=SCENE_SEED, max_points=None, pos=None, yaw_deg=None):
    """Return float32 (N,3) points in the SENSOR frame, in firing order (azimuth-major).  Default pose of frame k: 500 mm * k
    along x, yaw 0.5 deg * k (the scene ends after ~150 such frames); `pos` / `yaw_deg` place the sensor elsewhere."""
    vert, steps, max_range = SENSORS[sensor]
    boxes, cyls = _scene(scene_seed)
    yaw = np.deg2rad(0.5 * frame if yaw_deg is None else yaw_deg)
    pos = np.array([500.0 * frame, 0.0, 0.0]) if pos is None else np.asarray(pos, dtype=np.float64)
    az = np.arange(steps) * (2 * np.pi / steps)
    el = np.deg2rad(vert)
    A, E = np.meshgrid(az, el, indexing="ij")
    ds = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], axis=-1).reshape(-1, 3)
    cy, sy = np.cos(yaw), np.sin(yaw)
    R = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1.0]])
    dw = ds @ R.T
    t = _raycast(pos, dw, boxes, cyls)
    rng = np.random.default_rng(20260118 + frame)
    t = t + rng.normal(0.0, noise_mm, t.shape)
    keep = np.isfinite(t) & (t > 500.0) & (t <= max_range)
    pts = (ds[keep] * t[keep, None]).astype(np.float32)
    if max_points is not None and pts.shape[0] > max_points:
        pts = pts[:max_points]
    return np.ascontiguousarray(pts)


def make_lasers(sensor="hdl32e", frame=0, start_deg=123.45, dropout=0.02, noise_mm=20.0, scene_seed=SCENE_SEED, firings=None):
    """One rotation of raw returns as the reference's capture class hands them to the preprocessor (velodyne::Laser,
    include/VelodyneCapture.h:43-51): azimuth / vertical in degrees, distance in 2 mm units (0 = no return), in FIRING order
    -- the rotation starts at start_deg and wraps through 0.  The scene of make_scan plus what the preprocessor is there to
    remove: a ground plane 2450 mm under the sensor (src/preprocess.cpp:80-82), the roof of the own car, random dropouts.
    Azimuth convention of src/preprocess.cpp:50-52: x = d cos(v) sin(az), y = d cos(v) cos(az)."""
    vert, steps, max_range = SENSORS[sensor]
    native = 360.0 / steps   # firings < a rotation: a partial sweep at the sensor's own azimuth step
    steps = firings or steps
    boxes, cyls = _scene(scene_seed)
    ground = -1900.0
    height = 2450.0
    o = np.array([500.0 * frame, 0.0, ground + height])
    roof = np.array([[o[0] - 800.0, o[0] + 800.0, -1700.0, 1200.0, o[2] - 900.0, o[2] - 550.0]])
    boxes = np.concatenate([boxes, roof])
    az_deg = np.round((start_deg + np.arange(steps) * native) % 360.0, 2) % 360.0   # centi-degrees, like the device
    A, E = np.meshgrid(np.deg2rad(az_deg), np.deg2rad(vert), indexing="ij")
    d = np.stack([np.cos(E) * np.sin(A), np.cos(E) * np.cos(A), np.sin(E)], axis=-1).reshape(-1, 3)
    t = _raycast(o, d, boxes, cyls)
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = (ground - o[2]) / d[:, 2]
    t = np.where((tg > 0) & (tg < t), tg, t)
    rng = np.random.default_rng(20260118 + 977 * frame)
    t = t + rng.normal(0.0, noise_mm, t.shape)
    ok = np.isfinite(t) & (t > 400.0) & (t <= min(max_range, 131000.0)) & (rng.random(t.shape) >= dropout)
    dist = np.where(ok, np.round(np.where(ok, t, 0.0) / 2.0), 0).astype(np.uint16)
    return dict(azimuth=np.repeat(az_deg, len(vert)).astype(np.float64), vertical=np.tile(np.asarray(vert, np.float64), steps),
                distance=dist, ring_deg=np.sort(np.asarray(vert, np.float64)))


def random_descriptors(n, seed=7, density=None):
    """(n,6) uint64 B-SHOT records: 352 random bits (i.i.d. p=0.5, or `density` bits set), pad 0."""
    rng = np.random.default_rng(seed)
    if density is None:
        w = rng.integers(0, 2 ** 63, size=(n, 6), dtype=np.uint64) * np.uint64(2) + \
            rng.integers(0, 2, size=(n, 6), dtype=np.uint64)
    else:
        bits = rng.random((n, 352)) < (density / 352.0)
        w = pack_bits(bits)
    w[:, 5] &= np.uint64(0xFFFFFFFF)
    return np.ascontiguousarray(w)


def pack_bits(bits):
    """(n,352) bool -> (n,6) uint64 in std::bitset<352> layout (bit i -> word i/64, bit i%64)."""
    n = bits.shape[0]
    b = np.zeros((n, 384), dtype=np.uint8)
    b[:, :352] = bits
    by = np.packbits(b.reshape(n, 48, 8), axis=-1, bitorder="little").reshape(n, 48)
    return np.ascontiguousarray(by).view(np.uint64).reshape(n, 6)


def unpack_bits(words):
    """(n,6) uint64 -> (n,352) bool"""
    by = np.ascontiguousarray(words).view(np.uint8).reshape(-1, 48)
    return np.unpackbits(by, axis=-1, bitorder="little")[:, :352].astype(bool)

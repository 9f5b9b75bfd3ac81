"""bshot_b200 -- Python (ctypes) binding of the B200-native B-SHOT front end.

The product is the C-ABI shared library `libbshot_b200.so` (include/bshot_b200.h) built from the
sm_100a CUDA sources in csrc/.  This module only marshals numpy / torch buffers to that ABI for the
tests and bench.py.  There is NO CPU fallback: if the library is missing or no B200 is present the
calls raise.  Reference-named host C++ shims live in host/ (see INTEGRATION.md).

The directory name `b-shot-slam_b200` is not an importable identifier; load this package with
`importlib` under the name `bshot_b200` (tests/conftest.py, bench.py and __graft_entry__.py do so).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BSHOT_LIB") or os.path.join(_HERE, "libbshot_b200.so")  # BSHOT_LIB: tuning builds (tools/build_variant.sh)
_LIB = None

SR_CV, SR_CVS, SR_CVSN = 0, 1, 2
NORMALS_REFERENCE, NORMALS_FULL = 0, 1

EXPORTS = [
    "bshot_params_default", "bshot_version", "bshot_last_error", "bshot_ctx_create",
    "bshot_ctx_destroy", "bshot_ctx_stream", "bshot_ctx_sync", "bshot_ctx_reset", "bshot_set_cloud",
    "bshot_detect_keypoints", "bshot_seg_ratio", "bshot_set_keypoints", "bshot_compute_normals",
    "bshot_query_normals", "bshot_set_normals", "bshot_compute_shot", "bshot_compute_lrf",
    "bshot_binarize", "bshot_compute_descriptors", "bshot_match", "bshot_match_mutual",
    "bshot_process_frame", "bshot_process_frame_dev", "bshot_fetch_frame", "bshot_ctx_enable_timing",
    "bshot_stage_times", "bshot_frame_counters", "bshot_map_reset", "bshot_map_append",
    "bshot_map_size", "bshot_match_shard_dev", "bshot_match_dev", "bshot_merge_cands_dev",
    "bshot_match_map", "bshot_reverse_owned_dev", "bshot_apply_rq_dev", "bshot_push_cands_dev",
    "bshot_reverse_owned_push_dev", "bshot_peer_barrier_dev", "bshot_peer_barrier_timeouts", "bshot_peer_barrier_reset", "bshot_launch_count", "bshot_popc_peak", "bshot_debug_counters", "bshot_map_append_dev", "bshot_comm_create", "bshot_comm_export", "bshot_comm_import",
    "bshot_comm_region", "bshot_comm_import_ptrs", "bshot_comm_destroy", "bshot_comm_check", "bshot_match_map_sharded_dev",
    "bshot_match_map_sharded", "bshot_gmap_create", "bshot_gmap_reset", "bshot_gmap_size", "bshot_gmap_add",
    "bshot_gmap_update_from_frame", "bshot_gmap_get_keypoints", "bshot_extract_frame", "bshot_match_frame_to_map", "bshot_frame_commit", "bshot_ransac",
    "bshot_preprocess", "bshot_preprocess_select", "bshot_extract_scan", "bshot_icp", "bshot_evaluate_estimation", "bshot_set_matcher",
]


class Params(C.Structure):
    _fields_ = [("kp_radius", C.c_float), ("kp_max_nn", C.c_int), ("sr_type", C.c_int),
                ("top_k", C.c_int), ("normal_radius", C.c_float), ("normal_max_nn", C.c_int),
                ("normals_mode", C.c_int), ("shot_radius", C.c_float)]


CAND_DTYPE = np.dtype([("k1", "<u8"), ("k2", "<u8"), ("rq", "<u4"), ("pad", "<u4")])
NONE_KEY = np.uint64(0xFFFFFFFFFFFFFFFF)


class BshotError(RuntimeError):
    pass


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into libbshot_b200.so (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], stdout=out, stderr=out)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise BshotError(f"{LIB_PATH} is missing: build it with `make -C b-shot-slam_b200/csrc` "
                             "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, sz, ci, cf = C.c_void_p, C.c_size_t, C.c_int, C.c_float
        L.bshot_last_error.restype = C.c_char_p
        L.bshot_version.restype = ci
        L.bshot_params_default.argtypes = [C.POINTER(Params)]
        L.bshot_ctx_create.argtypes = [C.POINTER(vp), ci, sz, sz, sz]
        L.bshot_ctx_destroy.argtypes = [vp]
        L.bshot_ctx_destroy.restype = None
        L.bshot_ctx_stream.argtypes = [vp]
        L.bshot_ctx_stream.restype = vp
        L.bshot_ctx_sync.argtypes = [vp]
        L.bshot_ctx_reset.argtypes = [vp]
        L.bshot_set_cloud.argtypes = [vp, vp, sz, sz]
        L.bshot_detect_keypoints.argtypes = [vp, cf, ci, ci, ci, vp, vp, vp, vp]
        L.bshot_seg_ratio.argtypes = [vp, cf, ci, ci, vp]
        L.bshot_set_keypoints.argtypes = [vp, vp, sz, sz]
        L.bshot_compute_normals.argtypes = [vp, ci, cf, ci, vp]
        L.bshot_query_normals.argtypes = [vp, vp, sz, cf, ci, vp]
        L.bshot_set_normals.argtypes = [vp, vp, sz]
        L.bshot_compute_shot.argtypes = [vp, cf, vp, vp, vp, vp, vp]
        L.bshot_compute_lrf.argtypes = [vp, cf, vp, vp]
        L.bshot_binarize.argtypes = [vp, vp, sz, sz, vp]
        L.bshot_compute_descriptors.argtypes = [vp, C.POINTER(Params), vp]
        L.bshot_match.argtypes = [vp, vp, sz, vp, sz, vp, vp, vp, vp, vp]
        L.bshot_match_mutual.argtypes = [vp, vp, sz, vp, sz, vp, vp, vp]
        L.bshot_process_frame.argtypes = [vp, C.POINTER(Params), vp, sz, sz, vp, vp, vp, vp, vp]
        L.bshot_process_frame_dev.argtypes = [vp, C.POINTER(Params), vp, sz, sz]
        L.bshot_fetch_frame.argtypes = [vp, ci, vp, vp, vp, vp, vp]
        L.bshot_ctx_enable_timing.argtypes = [vp, ci]
        L.bshot_stage_times.argtypes = [vp, C.POINTER(C.c_float * 8)]
        L.bshot_frame_counters.argtypes = [vp, C.POINTER(C.c_ulonglong * 4)]
        L.bshot_map_reset.argtypes = [vp]
        L.bshot_map_append.argtypes = [vp, vp, sz]
        L.bshot_map_size.argtypes = [vp, C.POINTER(sz)]
        L.bshot_map_append_dev.argtypes = [vp, vp, sz]
        L.bshot_gmap_create.argtypes = [vp, sz, sz]
        L.bshot_gmap_reset.argtypes = [vp]
        L.bshot_gmap_size.argtypes = [vp, C.POINTER(sz), C.POINTER(sz)]
        L.bshot_gmap_add.argtypes = [vp, vp, vp, vp, sz, vp]
        L.bshot_gmap_update_from_frame.argtypes = [vp, vp]
        L.bshot_gmap_get_keypoints.argtypes = [vp, vp, cf, vp, vp, sz, C.POINTER(sz)]
        L.bshot_extract_frame.argtypes = [vp, C.POINTER(Params), vp, sz, sz, vp, vp, vp, vp, vp]
        L.bshot_match_frame_to_map.argtypes = [vp, vp, cf, vp, vp, vp, C.POINTER(sz), vp, sz]
        L.bshot_frame_commit.argtypes = [vp]
        L.bshot_ransac.argtypes = [vp, vp, sz, vp, sz, vp, sz, ci, cf, vp, vp, vp, vp]
        L.bshot_icp.argtypes = [vp, vp, sz, vp, sz, vp, ci, vp, vp, vp, vp]
        L.bshot_evaluate_estimation.argtypes = [vp, vp, vp, ci, vp, sz, vp, sz, ci, vp, vp, vp, vp, vp]
        L.bshot_preprocess_select.argtypes = [vp, vp, vp, vp, sz, vp, sz, C.c_double, C.c_double, vp, sz, ci, ci, vp, sz, C.POINTER(sz)]
        L.bshot_extract_scan.argtypes = [vp, C.POINTER(Params), vp, vp, vp, sz, vp, sz, C.c_double, C.c_double, vp, sz, C.POINTER(sz), vp, vp, vp, vp, vp]
        L.bshot_preprocess.argtypes = [vp, vp, vp, vp, sz, vp, sz, C.c_double, C.c_double, vp, sz, C.POINTER(sz)]
        L.bshot_comm_create.argtypes = [vp, ci, ci, sz]
        L.bshot_comm_export.argtypes = [vp, vp]
        L.bshot_comm_import.argtypes = [vp, vp]
        L.bshot_comm_region.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
        L.bshot_comm_import_ptrs.argtypes = [vp, C.POINTER(vp)]
        L.bshot_comm_destroy.argtypes = [vp]
        L.bshot_comm_check.argtypes = [vp]
        L.bshot_match_map_sharded_dev.argtypes = [vp, vp, sz, C.c_uint64, vp]
        L.bshot_match_map_sharded.argtypes = [vp, vp, sz, C.c_uint64, vp]
        L.bshot_match_shard_dev.argtypes = [vp, vp, sz, C.c_uint64, ci, vp]
        L.bshot_match_dev.argtypes = [vp, vp, sz, vp, sz, C.c_uint64, ci, vp]
        L.bshot_merge_cands_dev.argtypes = [vp, vp, sz, sz, vp]
        L.bshot_match_map.argtypes = [vp, vp, sz, C.c_uint64, vp]
        L.bshot_reverse_owned_dev.argtypes = [vp, vp, sz, C.c_uint64, vp, vp]
        L.bshot_apply_rq_dev.argtypes = [vp, vp, vp, sz]
        L.bshot_push_cands_dev.argtypes = [vp, vp, sz, vp, ci, ci]
        L.bshot_peer_barrier_dev.argtypes = [vp, vp, ci, ci]
        L.bshot_peer_barrier_timeouts.argtypes = [vp, C.POINTER(C.c_uint)]
        L.bshot_peer_barrier_reset.argtypes = [vp]
        L.bshot_reverse_owned_push_dev.argtypes = [vp, vp, sz, C.c_uint64, vp, vp, ci, ci]
        L.bshot_launch_count.argtypes = [vp]
        L.bshot_launch_count.restype = C.c_ulonglong
        L.bshot_popc_peak.argtypes = [vp, C.POINTER(C.c_double)]
        L.bshot_set_matcher.argtypes = [vp, ci]
        L.bshot_debug_counters.argtypes = [vp, C.POINTER(C.c_ulonglong * 8)]
        _LIB = L
    return _LIB


def _chk(rc):
    if rc != 0:
        raise BshotError(f"bshot error {rc}: {lib().bshot_last_error().decode()}")


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def default_params(**kw):
    p = Params()
    lib().bshot_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def unpack_cands(cand):
    """structured candidate records -> dict of int arrays (idx/dist, -1 = none)"""
    k1, k2 = cand["k1"], cand["k2"]
    h1, h2 = k1 != NONE_KEY, k2 != NONE_KEY
    out = dict(
        idx1=np.where(h1, (k1 & np.uint64(0xFFFFFFFF)).astype(np.int64), -1),
        dist1=np.where(h1, (k1 >> np.uint64(32)).astype(np.int64), -1),
        idx2=np.where(h2, (k2 & np.uint64(0xFFFFFFFF)).astype(np.int64), -1),
        dist2=np.where(h2, (k2 >> np.uint64(32)).astype(np.int64), -1),
        rq=np.where(cand["rq"] != 0xFFFFFFFF, cand["rq"].astype(np.int64), -1))
    return out


class Context:
    """One bshot_ctx: device buffers + one stream (mirrors the `bshot cb` member of LidarOdometry,
    include/lidar_odometry.h:57)."""

    def __init__(self, device=0, max_points=131072, max_keypoints=16384, max_targets=0):
        h = C.c_void_p()
        _chk(lib().bshot_ctx_create(C.byref(h), device, max_points, max_keypoints, max_targets))
        self.h = h
        self.max_points, self.max_keypoints = max_points, max_keypoints
        self.n_points = 0
        self.n_kp = 0

    def close(self):
        if getattr(self, "h", None):
            lib().bshot_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def stream(self):
        return lib().bshot_ctx_stream(self.h)

    def sync(self):
        _chk(lib().bshot_ctx_sync(self.h))

    def reset(self):
        _chk(lib().bshot_ctx_reset(self.h))

    def launch_count(self):
        return int(lib().bshot_launch_count(self.h))

    def set_matcher(self, kind):
        """-1 by problem size (default), 0 XOR + POPC, 1 tensor cores, 2 / 3 pipelined tensor-core kernel"""
        _chk(lib().bshot_set_matcher(self.h, int(kind)))

    def popc_peak(self):
        v = C.c_double()
        _chk(lib().bshot_popc_peak(self.h, C.byref(v)))
        return v.value

    # a1
    def set_cloud(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        assert xyz.ndim == 2 and xyz.shape[1] in (3, 4)
        _chk(lib().bshot_set_cloud(self.h, _p(xyz), xyz.shape[0], xyz.shape[1] * 4))
        self.n_points = xyz.shape[0]

    # a2 / a3
    def seg_ratio(self, radius=3000.0, max_nn=300, sr_type=SR_CV):
        out = np.empty(self.n_points, np.float32)
        _chk(lib().bshot_seg_ratio(self.h, radius, max_nn, sr_type, _p(out)))
        return out

    def detect_keypoints(self, radius=3000.0, max_nn=300, sr_type=SR_CV, top_k=600):
        idx = np.empty(top_k, np.int32)
        ratio = np.empty(top_k, np.float32)
        xyz = np.empty((top_k, 3), np.float32)
        cnt = C.c_int()
        _chk(lib().bshot_detect_keypoints(self.h, radius, max_nn, sr_type, top_k, _p(idx), _p(ratio),
                                          _p(xyz), C.byref(cnt)))
        k = cnt.value
        self.n_kp = k
        return idx[:k].copy(), ratio[:k].copy(), xyz[:k].copy()

    def set_keypoints(self, kp):
        kp = np.ascontiguousarray(kp, dtype=np.float32).reshape(-1, 3)
        _chk(lib().bshot_set_keypoints(self.h, _p(kp), kp.shape[0], 12))
        self.n_kp = kp.shape[0]

    # a4
    def compute_normals(self, mode=NORMALS_REFERENCE, radius=3000.0, max_nn=300):
        out = np.empty((self.n_points, 4), np.float32)
        _chk(lib().bshot_compute_normals(self.h, mode, radius, max_nn, _p(out)))
        return out

    def query_normals(self, q, radius=3000.0, max_nn=300):
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, 3)
        out = np.empty((q.shape[0], 4), np.float32)
        _chk(lib().bshot_query_normals(self.h, _p(q), q.shape[0], radius, max_nn, _p(out)))
        return out

    def set_normals(self, normals4):
        normals4 = np.ascontiguousarray(normals4, dtype=np.float32).reshape(-1, 4)
        _chk(lib().bshot_set_normals(self.h, _p(normals4), normals4.shape[0]))

    # a5..a7
    def compute_lrf(self, radius=3000.0):
        rf = np.empty((self.n_kp, 9), np.float32)
        nn = np.empty(self.n_kp, np.int32)
        _chk(lib().bshot_compute_lrf(self.h, radius, _p(rf), _p(nn)))
        return rf, nn

    def compute_shot(self, radius=3000.0, want_shot=True):
        k = self.n_kp
        bits = np.empty((k, 6), np.uint64)
        shot = np.empty((k, 352), np.float32) if want_shot else None
        rf = np.empty((k, 9), np.float32)
        nn = np.empty(k, np.int32)
        tot = C.c_longlong()
        _chk(lib().bshot_compute_shot(self.h, radius, _p(bits), _p(shot), _p(rf), _p(nn), C.byref(tot)))
        return dict(bits=bits, shot=shot, rf=rf, nn=nn, sum_neighbours=tot.value)

    def binarize(self, shot, stride_floats=None):
        shot = np.ascontiguousarray(shot, dtype=np.float32)
        if stride_floats is None:
            shot = shot.reshape(-1, 352)
            stride_floats = 352
        k = shot.size // stride_floats
        bits = np.empty((k, 6), np.uint64)
        _chk(lib().bshot_binarize(self.h, _p(shot), k, stride_floats, _p(bits)))
        return bits

    def compute_descriptors(self, params=None):
        bits = np.empty((self.n_kp, 6), np.uint64)
        _chk(lib().bshot_compute_descriptors(self.h, C.byref(params) if params else None, _p(bits)))
        return bits

    # a10 / a11
    def match(self, q, t, want_right=True):
        q = np.ascontiguousarray(q, dtype=np.uint64).reshape(-1, 6)
        t = np.ascontiguousarray(t, dtype=np.uint64).reshape(-1, 6)
        nq, nt = q.shape[0], t.shape[0]
        li, ld, li2, ld2 = (np.empty(nq, np.int32) for _ in range(4))
        ri = np.empty(nt, np.int32) if want_right else None
        _chk(lib().bshot_match(self.h, _p(q), nq, _p(t), nt, _p(li), _p(ld), _p(li2), _p(ld2), _p(ri)))
        return dict(left_idx=li, left_dist=ld, left_idx2=li2, left_dist2=ld2, right_idx=ri)

    def match_mutual(self, q, t):
        q = np.ascontiguousarray(q, dtype=np.uint64).reshape(-1, 6)
        t = np.ascontiguousarray(t, dtype=np.uint64).reshape(-1, 6)
        nq = q.shape[0]
        pairs = np.empty((max(nq, 1), 2), np.int32)
        dist = np.empty(max(nq, 1), np.int32)
        cnt = C.c_int()
        _chk(lib().bshot_match_mutual(self.h, _p(q), nq, _p(t), t.shape[0], _p(pairs), _p(dist), C.byref(cnt)))
        return pairs[:cnt.value].copy(), dist[:cnt.value].copy()

    # whole frame
    def process_frame(self, xyz, params):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        k = params.top_k
        kp_idx = np.empty(k, np.int32)
        bits = np.empty((k, 6), np.uint64)
        pairs = np.empty((k, 2), np.int32)
        nk, npairs = C.c_int(), C.c_int()
        _chk(lib().bshot_process_frame(self.h, C.byref(params), _p(xyz), xyz.shape[0], xyz.shape[1] * 4,
                                       _p(kp_idx), _p(bits), C.byref(nk), _p(pairs), C.byref(npairs)))
        self.n_points = xyz.shape[0]
        self.n_kp = nk.value
        return dict(kp_idx=kp_idx[:nk.value].copy(), bits=bits[:nk.value].copy(),
                    pairs=pairs[:npairs.value].copy())

    def process_frame_raw(self, xyz_ptr, n, stride_bytes, params, kp_idx_ptr, bits_ptr, pairs_ptr):
        """pointer-level variant for bench.py (pinned host buffers, no numpy allocation)"""
        nk, npairs = C.c_int(), C.c_int()
        _chk(lib().bshot_process_frame(self.h, C.byref(params), xyz_ptr, n, stride_bytes, kp_idx_ptr,
                                       bits_ptr, C.byref(nk), pairs_ptr, C.byref(npairs)))
        return nk.value, npairs.value

    def process_frame_dev(self, d_xyz_ptr, n, stride_bytes, params):
        """device-resident cloud (raw device pointer), asynchronous on the context stream"""
        _chk(lib().bshot_process_frame_dev(self.h, C.byref(params), d_xyz_ptr, n, stride_bytes))

    def fetch_frame(self, top_k):
        kp_idx = np.empty(top_k, np.int32)
        bits = np.empty((top_k, 6), np.uint64)
        pairs = np.empty((top_k, 2), np.int32)
        nk, npairs = C.c_int(), C.c_int()
        _chk(lib().bshot_fetch_frame(self.h, top_k, _p(kp_idx), _p(bits), C.byref(nk), _p(pairs), C.byref(npairs)))
        self.n_kp = nk.value
        return dict(kp_idx=kp_idx[:nk.value].copy(), bits=bits[:nk.value].copy(), pairs=pairs[:npairs.value].copy())

    def enable_timing(self, on=True):
        _chk(lib().bshot_ctx_enable_timing(self.h, int(on)))

    def stage_times(self):
        a = (C.c_float * 8)()
        _chk(lib().bshot_stage_times(self.h, C.byref(a)))
        names = ["voxel_build", "seg_ratio", "topk", "normals", "shot_bshot", "match", "frame"]
        return {k: float(a[i]) for i, k in enumerate(names)}

    def frame_counters(self):
        a = (C.c_ulonglong * 4)()
        _chk(lib().bshot_frame_counters(self.h, C.byref(a)))
        return dict(detector_neighbours=int(a[0]), normals_neighbours=int(a[1]), shot_neighbours=int(a[2]),
                    keypoints=int(a[3]))

    def debug_counters(self):
        a = (C.c_ulonglong * 8)()
        _chk(lib().bshot_debug_counters(self.h, C.byref(a)))
        names = ["detector_neighbours", "normals_neighbours", "tiles_staged", "tile_points_swept", "query_attempts",
                 "unresolved_attempts", "fallback_queries", "blocks"]
        return {k: int(a[i]) for i, k in enumerate(names)}

    # sharded map
    def map_reset(self):
        _chk(lib().bshot_map_reset(self.h))

    def map_append(self, desc):
        desc = np.ascontiguousarray(desc, dtype=np.uint64).reshape(-1, 6)
        _chk(lib().bshot_map_append(self.h, _p(desc), desc.shape[0]))

    def map_append_dev(self, d_desc_ptr, n):
        _chk(lib().bshot_map_append_dev(self.h, d_desc_ptr, n))

    # GPU-resident global map + frame-to-map flow
    def gmap_create(self, max_entries, max_blocks=4096):
        _chk(lib().bshot_gmap_create(self.h, max_entries, max_blocks))

    def gmap_reset(self):
        _chk(lib().bshot_gmap_reset(self.h))

    def gmap_size(self):
        n, d = C.c_size_t(), C.c_size_t()
        _chk(lib().bshot_gmap_size(self.h, C.byref(n), C.byref(d)))
        return n.value, d.value

    def gmap_add(self, xyz, ratio, desc, pose=None):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        ratio = np.ascontiguousarray(ratio, dtype=np.float32).reshape(-1)
        desc = np.ascontiguousarray(desc, dtype=np.uint64).reshape(-1, 6)
        pose = None if pose is None else np.ascontiguousarray(pose, dtype=np.float32).reshape(12)
        _chk(lib().bshot_gmap_add(self.h, _p(xyz), _p(ratio), _p(desc), xyz.shape[0], _p(pose)))

    def gmap_update_from_frame(self, pose=None):
        pose = None if pose is None else np.ascontiguousarray(pose, dtype=np.float32).reshape(12)
        _chk(lib().bshot_gmap_update_from_frame(self.h, _p(pose)))

    def gmap_get_keypoints(self, pos, rng, cap=1 << 20):
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(3)
        n = C.c_size_t()
        _chk(lib().bshot_gmap_get_keypoints(self.h, _p(pos), rng, None, None, 0, C.byref(n)))
        xyz = np.empty((n.value, 3), np.float32)
        desc = np.empty((n.value, 6), np.uint64)
        _chk(lib().bshot_gmap_get_keypoints(self.h, _p(pos), rng, _p(xyz), _p(desc), n.value, C.byref(n)))
        return xyz, desc

    def extract_frame(self, xyz, params):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        k = params.top_k
        idx, kp, ratio, bits = np.empty(k, np.int32), np.empty((k, 3), np.float32), np.empty(k, np.float32), np.empty((k, 6), np.uint64)
        nk = C.c_int()
        _chk(lib().bshot_extract_frame(self.h, C.byref(params), _p(xyz), xyz.shape[0], xyz.shape[1] * 4, _p(idx), _p(kp), _p(ratio),
                                       _p(bits), C.byref(nk)))
        n = nk.value
        self.n_points, self.n_kp = xyz.shape[0], n
        return dict(kp_idx=idx[:n].copy(), kp_xyz=kp[:n].copy(), seg_ratio=ratio[:n].copy(), bits=bits[:n].copy())

    def match_frame_to_map(self, ref_pos, rng=100000.0, ref_pose=None, target_cap=1 << 20):
        ref_pos = np.ascontiguousarray(ref_pos, dtype=np.float32).reshape(3)
        ref_pose = None if ref_pose is None else np.ascontiguousarray(ref_pose, dtype=np.float32).reshape(12)
        pairs = np.empty((max(self.max_keypoints, 1), 2), np.int32)
        npairs, nt = C.c_int(), C.c_size_t()
        txyz = np.empty((target_cap, 3), np.float32)
        _chk(lib().bshot_match_frame_to_map(self.h, _p(ref_pos), rng, _p(ref_pose), _p(pairs), C.byref(npairs), C.byref(nt), _p(txyz), target_cap))
        return dict(pairs=pairs[:npairs.value].copy(), n_targets=nt.value, target_xyz=txyz[:nt.value].copy())

    def ransac(self, src_xyz, tgt_xyz, pairs, max_iterations=2000, threshold=1500.0):
        src_xyz = np.ascontiguousarray(src_xyz, dtype=np.float32).reshape(-1, 3)
        tgt_xyz = np.ascontiguousarray(tgt_xyz, dtype=np.float32).reshape(-1, 3)
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        out = np.empty((max(len(pairs), 1), 2), np.int32)
        T = np.empty((4, 4), np.float32)
        n, it = C.c_int(), C.c_int()
        _chk(lib().bshot_ransac(self.h, _p(src_xyz), src_xyz.shape[0], _p(tgt_xyz), tgt_xyz.shape[0], _p(pairs), pairs.shape[0],
                                max_iterations, threshold, _p(out), C.byref(n), _p(T), C.byref(it)))
        return dict(pairs=out[:n.value].copy(), transform=T, iterations=it.value)

    def icp(self, src_xyz, tgt_xyz, pre=None, max_iterations=10):
        """pcl::IterativeClosestPoint with PCL's defaults (src/lidar_odometry.cpp:283-291) on src moved by `pre`"""
        src_xyz = np.ascontiguousarray(src_xyz, dtype=np.float32).reshape(-1, 3)
        tgt_xyz = np.ascontiguousarray(tgt_xyz, dtype=np.float32).reshape(-1, 3)
        pre = None if pre is None else np.ascontiguousarray(pre, dtype=np.float32).reshape(16)
        T = np.empty((4, 4), np.float32)
        it, st, mse = C.c_int(), C.c_int(), C.c_double()
        _chk(lib().bshot_icp(self.h, _p(src_xyz), src_xyz.shape[0], _p(tgt_xyz), tgt_xyz.shape[0], _p(pre), max_iterations, _p(T), C.byref(it),
                             C.byref(st), C.byref(mse)))
        return dict(transform=T, iterations=it.value, state=st.value, mse=mse.value)

    def evaluate_estimation(self, T_ransac, T_ref, n_corr, src_kp, tgt_kp, run_icp=True):
        """LidarOdometry::evaluateEstimation (src/lidar_odometry.cpp:267-296)"""
        T_ransac = np.ascontiguousarray(T_ransac, dtype=np.float32).reshape(16)
        T_ref = np.ascontiguousarray(T_ref, dtype=np.float32).reshape(16)
        src_kp = np.ascontiguousarray(src_kp, dtype=np.float32).reshape(-1, 3)
        tgt_kp = np.ascontiguousarray(tgt_kp, dtype=np.float32).reshape(-1, 3)
        T = np.empty((4, 4), np.float32)
        upd, it, h, t = C.c_int(), C.c_int(), C.c_float(), C.c_float()
        _chk(lib().bshot_evaluate_estimation(self.h, _p(T_ransac), _p(T_ref), n_corr, _p(src_kp), src_kp.shape[0], _p(tgt_kp), tgt_kp.shape[0],
                                             1 if run_icp else 0, _p(T), C.byref(upd), C.byref(h), C.byref(t), C.byref(it)))
        return dict(T_best=T, should_update_map=bool(upd.value), h_diff=h.value, t_diff=t.value, icp_iterations=it.value)

    def preprocess(self, azimuth_deg, vertical_deg, distance, ring_deg, vert_init=-0.6, lowpt_th=-1950.0, select=None, save_selected=True):
        """Preprocessor::run (src/preprocess.cpp:213-223) on one rotation of returns sorted by azimuth -> (m, 3) float32 mm."""
        az = np.ascontiguousarray(azimuth_deg, dtype=np.float64)
        ve = np.ascontiguousarray(vertical_deg, dtype=np.float64)
        di = np.ascontiguousarray(distance, dtype=np.uint16)
        ring = np.ascontiguousarray(ring_deg, dtype=np.float64)
        assert az.shape == ve.shape == di.shape
        out = np.empty((max(az.size, 1), 3), np.float32)
        n = C.c_size_t()
        sel = None if select is None else np.ascontiguousarray(select, dtype=np.int32)
        _chk(lib().bshot_preprocess_select(self.h, _p(az), _p(ve), _p(di), az.size, _p(ring), ring.size, vert_init, lowpt_th, _p(sel),
                                           0 if sel is None else sel.size, 0 if sel is None else 1, 1 if save_selected else 0, _p(out), out.shape[0], C.byref(n)))
        return out[:n.value].copy()

    def extract_scan(self, azimuth_deg, vertical_deg, distance, ring_deg, params, vert_init=-0.6, lowpt_th=-1950.0, want_cloud=True):
        """lasers -> preprocessed cloud (device resident) -> keypoints + descriptors in one call"""
        az = np.ascontiguousarray(azimuth_deg, dtype=np.float64)
        ve = np.ascontiguousarray(vertical_deg, dtype=np.float64)
        di = np.ascontiguousarray(distance, dtype=np.uint16)
        ring = np.ascontiguousarray(ring_deg, dtype=np.float64)
        k = params.top_k
        cloud = np.empty((max(az.size, 1), 3), np.float32) if want_cloud else None
        idx, kp, ratio, bits = np.empty(k, np.int32), np.empty((k, 3), np.float32), np.empty(k, np.float32), np.empty((k, 6), np.uint64)
        npts, nk = C.c_size_t(), C.c_int()
        _chk(lib().bshot_extract_scan(self.h, C.byref(params), _p(az), _p(ve), _p(di), az.size, _p(ring), ring.size, vert_init, lowpt_th, _p(cloud),
                                      0 if cloud is None else cloud.shape[0], C.byref(npts), _p(idx), _p(kp), _p(ratio), _p(bits), C.byref(nk)))
        n = nk.value
        self.n_points, self.n_kp = npts.value, n
        return dict(cloud=None if cloud is None else cloud[:npts.value].copy(), n_points=npts.value, kp_idx=idx[:n].copy(), kp_xyz=kp[:n].copy(),
                    seg_ratio=ratio[:n].copy(), bits=bits[:n].copy())

    def frame_commit(self):
        _chk(lib().bshot_frame_commit(self.h))

    # multi-rank exchange behind the C ABI (CUDA IPC peer memory)
    def comm_create(self, rank, nranks, max_queries):
        _chk(lib().bshot_comm_create(self.h, rank, nranks, max_queries))

    def comm_export(self):
        h = np.zeros(64, np.uint8)
        _chk(lib().bshot_comm_export(self.h, _p(h)))
        return h

    def comm_import(self, handles):
        handles = np.ascontiguousarray(handles, dtype=np.uint8).reshape(-1, 64)
        _chk(lib().bshot_comm_import(self.h, _p(handles)))

    def comm_region(self):
        ptr, n = C.c_void_p(), C.c_size_t()
        _chk(lib().bshot_comm_region(self.h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def comm_import_ptrs(self, ptrs):
        arr = (C.c_void_p * len(ptrs))(*[C.c_void_p(p) for p in ptrs])
        _chk(lib().bshot_comm_import_ptrs(self.h, arr))

    def comm_destroy(self):
        _chk(lib().bshot_comm_destroy(self.h))

    def comm_check(self):
        _chk(lib().bshot_comm_check(self.h))

    def match_map_sharded_dev(self, d_q_ptr, nq, global_base, d_cand_ptr):
        _chk(lib().bshot_match_map_sharded_dev(self.h, d_q_ptr, nq, global_base, d_cand_ptr))

    def match_map_sharded(self, q, global_base=0):
        q = np.ascontiguousarray(q, dtype=np.uint64).reshape(-1, 6)
        cand = np.empty(q.shape[0], CAND_DTYPE)
        _chk(lib().bshot_match_map_sharded(self.h, _p(q), q.shape[0], global_base, _p(cand)))
        return cand

    def map_size(self):
        n = C.c_size_t()
        _chk(lib().bshot_map_size(self.h, C.byref(n)))
        return n.value

    def match_map(self, q, global_base=0):
        q = np.ascontiguousarray(q, dtype=np.uint64).reshape(-1, 6)
        cand = np.empty(q.shape[0], CAND_DTYPE)
        _chk(lib().bshot_match_map(self.h, _p(q), q.shape[0], global_base, _p(cand)))
        return cand

    def match_shard_dev(self, d_q_ptr, nq, global_base, with_rq, d_cand_ptr):
        _chk(lib().bshot_match_shard_dev(self.h, d_q_ptr, nq, global_base, int(with_rq), d_cand_ptr))

    def match_dev(self, d_q_ptr, nq, d_t_ptr, nt, global_base, with_rq, d_cand_ptr):
        _chk(lib().bshot_match_dev(self.h, d_q_ptr, nq, d_t_ptr, nt, global_base, int(with_rq), d_cand_ptr))

    def reverse_owned_dev(self, d_q_ptr, nq, global_base, d_merged_ptr, d_rq_ptr):
        _chk(lib().bshot_reverse_owned_dev(self.h, d_q_ptr, nq, global_base, d_merged_ptr, d_rq_ptr))

    def apply_rq_dev(self, d_cands_ptr, d_rq_ptr, nq):
        _chk(lib().bshot_apply_rq_dev(self.h, d_cands_ptr, d_rq_ptr, nq))

    # peer-memory exchange (symmetric buffers): stores into every rank's buffers, the caller adds the barriers
    def push_cands_dev(self, d_cands_ptr, nq, d_peer_ptrs, nranks, rank):
        _chk(lib().bshot_push_cands_dev(self.h, d_cands_ptr, nq, d_peer_ptrs, nranks, rank))

    def peer_barrier_dev(self, d_peer_flag_ptrs, nranks, rank):
        _chk(lib().bshot_peer_barrier_dev(self.h, d_peer_flag_ptrs, nranks, rank))

    def peer_barrier_reset(self):
        _chk(lib().bshot_peer_barrier_reset(self.h))

    def peer_barrier_timeouts(self):
        e = C.c_uint()
        _chk(lib().bshot_peer_barrier_timeouts(self.h, C.byref(e)))
        return int(e.value)

    def reverse_owned_push_dev(self, d_q_ptr, nq, global_base, d_merged_ptr, d_peer_rq_ptrs, nranks, rank):
        _chk(lib().bshot_reverse_owned_push_dev(self.h, d_q_ptr, nq, global_base, d_merged_ptr, d_peer_rq_ptrs, nranks, rank))

    def merge_cands_dev(self, d_cands_ptr, nranks, nq, d_out_ptr):
        _chk(lib().bshot_merge_cands_dev(self.h, d_cands_ptr, nranks, nq, d_out_ptr))

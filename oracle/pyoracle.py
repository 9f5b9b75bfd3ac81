"""ctypes binding of the CPU oracle (oracle/bshot_oracle.{h,cpp}).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
Parity pin: see the header of oracle/bshot_oracle.h (reference-owned arithmetic pinned to oracle/_ref,
PCL-owned arithmetic unpinned).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SR_CV, SR_CVS, SR_CVSN = 0, 1, 2
TIE_STDSORT, TIE_DETERMINISTIC = 0, 1
MODE_REFERENCE, MODE_FULL = 0, 1


def build(force=False):
    so = os.path.join(_HERE, "libbshot_oracle.so")
    src = os.path.join(_HERE, "bshot_oracle.cpp")
    hdr = os.path.join(_HERE, "bshot_oracle.h")
    stale = (not os.path.exists(so)) or any(
        os.path.getmtime(p) > os.path.getmtime(so) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libbshot_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        fp = C.POINTER(C.c_float)
        ip = C.POINTER(C.c_int)
        up = C.POINTER(C.c_uint64)
        L.orc_cloud_create.restype = C.c_void_p
        L.orc_cloud_create.argtypes = [fp, C.c_size_t, C.c_size_t]
        L.orc_cloud_destroy.argtypes = [C.c_void_p]
        L.orc_cloud_size.restype = C.c_size_t
        L.orc_cloud_size.argtypes = [C.c_void_p]
        L.orc_radius_search.restype = C.c_int
        L.orc_radius_search.argtypes = [C.c_void_p, fp, C.c_float, C.c_int, ip, fp, C.c_int]
        L.orc_seg_ratio.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int, fp, C.c_int]
        L.orc_select_keypoints.restype = C.c_int
        L.orc_select_keypoints.argtypes = [fp, C.c_size_t, C.c_int, C.c_int, ip, fp]
        L.orc_normals.argtypes = [C.c_void_p, fp, C.c_size_t, C.c_float, C.c_int, fp, C.c_int]
        L.orc_lrf.argtypes = [C.c_void_p, fp, C.c_size_t, C.c_float, fp, ip, C.c_int]
        L.orc_shot.restype = C.c_longlong
        L.orc_shot.argtypes = [C.c_void_p, fp, C.c_size_t, C.c_float, fp, fp, fp, fp, ip, C.c_int]
        L.orc_bshot.argtypes = [fp, C.c_size_t, up]
        L.orc_match.argtypes = [up, C.c_size_t, up, C.c_size_t, ip, ip, ip, ip, ip, C.c_int]
        L.orc_mutual.restype = C.c_int
        L.orc_mutual.argtypes = [ip, C.c_size_t, ip, ip]
        L.orc_compute_descriptors.restype = C.c_longlong
        L.orc_compute_descriptors.argtypes = [C.c_void_p, fp, C.c_size_t, C.c_float, C.c_int,
                                              C.c_int, up, fp, fp, fp, C.c_int]
        L.orc_map_create.restype = C.c_void_p
        L.orc_map_destroy.argtypes = [C.c_void_p]
        L.orc_map_size.restype = C.c_size_t
        L.orc_map_size.argtypes = [C.c_void_p]
        L.orc_map_add.argtypes = [C.c_void_p, fp, fp, up, C.c_size_t, fp]
        L.orc_map_get.restype = C.c_size_t
        L.orc_map_get.argtypes = [C.c_void_p, fp, C.c_float, fp, up, C.c_size_t]
        L.orc_ransac.restype = C.c_int
        L.orc_ransac.argtypes = [fp, fp, ip, C.c_size_t, C.c_int, C.c_double, ip, fp, ip]
        L.orc_eigh3.argtypes = [C.POINTER(C.c_double)] * 3
        L.orc_eigen33_smallest.argtypes = [fp, fp, fp]
        _LIB = L
    return _LIB


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _u(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


class Cloud:
    """surface cloud (N,3) float32 in mm + search grid"""

    def __init__(self, xyz):
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        self.h = lib().orc_cloud_create(_f(self.xyz), self.xyz.shape[0], 3)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_cloud_destroy(self.h)
            self.h = None

    def __len__(self):
        return self.xyz.shape[0]

    def radius_search(self, q, radius, max_nn=0):
        q = np.ascontiguousarray(q, dtype=np.float32)
        cap = len(self)
        idx = np.empty(cap, np.int32)
        sqd = np.empty(cap, np.float32)
        n = lib().orc_radius_search(self.h, _f(q), radius, max_nn, _i(idx), _f(sqd), cap)
        return idx[:n].copy(), sqd[:n].copy()

    def seg_ratio(self, radius=3000.0, max_nn=300, sr_type=SR_CV, threads=0):
        out = np.empty(len(self), np.float32)
        lib().orc_seg_ratio(self.h, radius, max_nn, sr_type, _f(out), threads)
        return out

    def normals(self, q, radius=3000.0, max_nn=300, threads=0):
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, 3)
        out = np.empty((q.shape[0], 4), np.float32)
        lib().orc_normals(self.h, _f(q), q.shape[0], radius, max_nn, _f(out), threads)
        return out

    def lrf(self, kp, radius=3000.0, threads=0):
        kp = np.ascontiguousarray(kp, dtype=np.float32).reshape(-1, 3)
        rf = np.empty((kp.shape[0], 9), np.float32)
        valid = np.empty(kp.shape[0], np.int32)
        lib().orc_lrf(self.h, _f(kp), kp.shape[0], radius, _f(rf), _i(valid), threads)
        return rf, valid

    def shot(self, kp, normals4, radius=3000.0, rf_in=None, threads=0):
        kp = np.ascontiguousarray(kp, dtype=np.float32).reshape(-1, 3)
        normals4 = np.ascontiguousarray(normals4, dtype=np.float32).reshape(len(self), 4)
        shot = np.empty((kp.shape[0], 352), np.float32)
        rf = np.empty((kp.shape[0], 9), np.float32)
        nn = np.empty(kp.shape[0], np.int32)
        rfi = None
        if rf_in is not None:
            rf_in = np.ascontiguousarray(rf_in, dtype=np.float32).reshape(-1, 9)
            rfi = _f(rf_in)
        total = lib().orc_shot(self.h, _f(kp), kp.shape[0], radius, _f(normals4), rfi, _f(shot),
                               _f(rf), _i(nn), threads)
        return shot, rf, nn, total

    def compute_descriptors(self, kp, radius=3000.0, max_nn=300, mode=MODE_REFERENCE, threads=0,
                            want_normals=False):
        kp = np.ascontiguousarray(kp, dtype=np.float32).reshape(-1, 3)
        k = kp.shape[0]
        bits = np.empty((k, 6), np.uint64)
        shot = np.empty((k, 352), np.float32)
        rf = np.empty((k, 9), np.float32)
        normals = np.empty((len(self), 4), np.float32) if want_normals else None
        total = lib().orc_compute_descriptors(
            self.h, _f(kp), k, radius, max_nn, mode, _u(bits), _f(shot), _f(rf),
            _f(normals) if want_normals else None, threads)
        return dict(bits=bits, shot=shot, rf=rf, normals=normals, sum_neighbours=total)


class Map:
    """the reference's global keypoint map (src/mymap.cpp, src/keypoint.cpp:23-32); insertion order inside a block"""

    def __init__(self):
        self.h = lib().orc_map_create()

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_map_destroy(self.h)
            self.h = None

    def __len__(self):
        return lib().orc_map_size(self.h)

    def add(self, xyz, ratio, desc, pose=None):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        ratio = np.ascontiguousarray(ratio, dtype=np.float32).reshape(-1)
        desc = np.ascontiguousarray(desc, dtype=np.uint64).reshape(-1, 6)
        pose = None if pose is None else np.ascontiguousarray(pose, dtype=np.float32).reshape(12)
        lib().orc_map_add(self.h, _f(xyz), _f(ratio), _u(desc), xyz.shape[0], None if pose is None else _f(pose))

    def get(self, pos, rng):
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(3)
        n = lib().orc_map_get(self.h, _f(pos), rng, None, None, 0)
        xyz = np.empty((n, 3), np.float32)
        desc = np.empty((n, 6), np.uint64)
        lib().orc_map_get(self.h, _f(pos), rng, _f(xyz), _u(desc), n)
        return xyz, desc


def ransac(src_xyz, tgt_xyz, pairs, max_iterations=2000, threshold=1500.0):
    """PCL 1.8 CorrespondenceRejectorSampleConsensus as the reference calls it (src/lidar_odometry.cpp:251-261)"""
    src_xyz = np.ascontiguousarray(src_xyz, dtype=np.float32).reshape(-1, 3)
    tgt_xyz = np.ascontiguousarray(tgt_xyz, dtype=np.float32).reshape(-1, 3)
    pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    out = np.empty((max(len(pairs), 1), 2), np.int32)
    T = np.empty((4, 4), np.float32)
    it = C.c_int()
    n = lib().orc_ransac(_f(src_xyz), _f(tgt_xyz), _i(pairs), pairs.shape[0], max_iterations, threshold, _i(out), _f(T), C.byref(it))
    return dict(pairs=out[:n].copy(), transform=T, iterations=it.value)


def icp(src_xyz, tgt_xyz, pre=None, max_iterations=10):
    """pcl::IterativeClosestPoint with PCL's defaults as src/lidar_odometry.cpp:283-291 uses it"""
    src_xyz = np.ascontiguousarray(src_xyz, dtype=np.float32).reshape(-1, 3)
    tgt_xyz = np.ascontiguousarray(tgt_xyz, dtype=np.float32).reshape(-1, 3)
    pre = None if pre is None else np.ascontiguousarray(pre, dtype=np.float32).reshape(16)
    T = np.empty((4, 4), np.float32)
    it, mse = C.c_int(), C.c_double()
    L = lib()
    L.orc_icp.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    st = L.orc_icp(src_xyz.ctypes.data, src_xyz.shape[0], tgt_xyz.ctypes.data, tgt_xyz.shape[0], None if pre is None else pre.ctypes.data,
                   max_iterations, T.ctypes.data, C.addressof(it), C.addressof(mse))
    return dict(transform=T, iterations=it.value, state=st, mse=mse.value)


def evaluate_estimation(T_ransac, T_ref, n_corr, src_kp, tgt_kp, run_icp=True):
    """LidarOdometry::evaluateEstimation (src/lidar_odometry.cpp:267-296)"""
    T_ransac = np.ascontiguousarray(T_ransac, dtype=np.float32).reshape(16)
    T_ref = np.ascontiguousarray(T_ref, dtype=np.float32).reshape(16)
    src_kp = np.ascontiguousarray(src_kp, dtype=np.float32).reshape(-1, 3)
    tgt_kp = np.ascontiguousarray(tgt_kp, dtype=np.float32).reshape(-1, 3)
    T = np.empty((4, 4), np.float32)
    h, t = C.c_float(), C.c_float()
    L = lib()
    L.orc_evaluate_estimation.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                          C.c_void_p, C.c_void_p]
    upd = L.orc_evaluate_estimation(T_ransac.ctypes.data, T_ref.ctypes.data, n_corr, src_kp.ctypes.data, src_kp.shape[0], tgt_kp.ctypes.data,
                                    tgt_kp.shape[0], 1 if run_icp else 0, T.ctypes.data, C.addressof(h), C.addressof(t))
    return dict(T_best=T, should_update_map=bool(upd), h_diff=h.value, t_diff=t.value)


def select_keypoints(ratio, top_k=600, tie_mode=TIE_DETERMINISTIC):
    ratio = np.ascontiguousarray(ratio, dtype=np.float32)
    idx = np.empty(max(top_k, 1), np.int32)
    rat = np.empty(max(top_k, 1), np.float32)
    n = lib().orc_select_keypoints(_f(ratio), ratio.shape[0], top_k, tie_mode, _i(idx), _f(rat))
    return idx[:n].copy(), rat[:n].copy()


def bshot(shot):
    shot = np.ascontiguousarray(shot, dtype=np.float32).reshape(-1, 352)
    bits = np.empty((shot.shape[0], 6), np.uint64)
    lib().orc_bshot(_f(shot), shot.shape[0], _u(bits))
    return bits


def match(q, t, want_right=True, threads=0):
    q = np.ascontiguousarray(q, dtype=np.uint64).reshape(-1, 6)
    t = np.ascontiguousarray(t, dtype=np.uint64).reshape(-1, 6)
    li = np.empty(q.shape[0], np.int32)
    ld = np.empty(q.shape[0], np.int32)
    li2 = np.empty(q.shape[0], np.int32)
    ld2 = np.empty(q.shape[0], np.int32)
    ri = np.empty(t.shape[0], np.int32) if want_right else None
    lib().orc_match(_u(q), q.shape[0], _u(t), t.shape[0], _i(li), _i(ld), _i(li2), _i(ld2),
                    _i(ri) if want_right else None, threads)
    return dict(left_idx=li, left_dist=ld, left_idx2=li2, left_dist2=ld2, right_idx=ri)


def mutual(left_idx, right_idx):
    left_idx = np.ascontiguousarray(left_idx, dtype=np.int32)
    right_idx = np.ascontiguousarray(right_idx, dtype=np.int32)
    pairs = np.empty((left_idx.shape[0], 2), np.int32)
    n = lib().orc_mutual(_i(left_idx), left_idx.shape[0], _i(right_idx), _i(pairs))
    return pairs[:n].copy()


def eigh3(m):
    m = np.ascontiguousarray(m, dtype=np.float64).reshape(9)
    w = np.empty(3, np.float64)
    v = np.empty(9, np.float64)
    dp = C.POINTER(C.c_double)
    lib().orc_eigh3(m.ctypes.data_as(dp), w.ctypes.data_as(dp), v.ctypes.data_as(dp))
    return w, v.reshape(3, 3)


def eigen33_smallest(m):
    m = np.ascontiguousarray(m, dtype=np.float32).reshape(9)
    ev = np.empty(1, np.float32)
    vec = np.empty(3, np.float32)
    lib().orc_eigen33_smallest(_f(m), _f(ev), _f(vec))
    return float(ev[0]), vec


# ---- oracle/_ref: the reference's own header compiled unchanged (oracle/ref_shim.cpp, oracle/Makefile) ---------
_REF = None


def ref_lib():
    """libbshot_ref.so (include/bshot_bits.h of the reference compiled against oracle/pcl_stub), or None when it
    has not been built (it can only be built where /root/reference exists; the built .so travels to the GPU box)."""
    global _REF
    if _REF is None:
        so = os.path.join(_HERE, "_ref", "libbshot_ref.so")
        if os.path.isdir("/root/reference/include"):
            subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
        if not os.path.exists(so):
            return None
        L = C.CDLL(so)
        fp, ip, up = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_uint64)
        L.ref_bshot_from_shot.argtypes = [fp, C.c_size_t, up]
        L.ref_minvect_int.restype = C.c_int
        L.ref_minvect_int.argtypes = [ip, C.c_int, ip]
        L.ref_feature_matching.restype = C.c_int
        L.ref_feature_matching.argtypes = [up, C.c_size_t, up, C.c_size_t, ip, ip, ip]
        L.ref_cb_create.restype = C.c_void_p
        L.ref_cb_destroy.argtypes = [C.c_void_p]
        L.ref_cb_compute_descriptors.argtypes = [C.c_void_p, fp, C.c_size_t, fp, C.c_size_t, C.c_float, up, fp, fp, fp]
        dp = C.POINTER(C.c_double)
        L.ref_preprocess.restype = C.c_size_t
        L.ref_preprocess_select.restype = C.c_size_t
        L.ref_preprocess_select.argtypes = [dp, dp, C.POINTER(C.c_ushort), C.c_size_t, dp, C.c_size_t, C.c_double, C.c_double, ip, C.c_size_t, C.c_int, C.c_int, fp, C.c_size_t]
        L.ref_preprocess.argtypes = [dp, dp, C.POINTER(C.c_ushort), C.c_size_t, dp, C.c_size_t, C.c_double, C.c_double, fp, C.c_size_t]
        _REF = L
    return _REF


def ref_bshot(shot):
    """bshot::compute_bshot_from_SHOT of the reference (include/bshot_bits.h:144-278), compiled unchanged"""
    shot = np.ascontiguousarray(shot, dtype=np.float32).reshape(-1, 352)
    bits = np.empty((shot.shape[0], 6), np.uint64)
    ref_lib().ref_bshot_from_shot(_f(shot), shot.shape[0], _u(bits))
    return bits


def ref_minvect(v):
    v = np.ascontiguousarray(v, dtype=np.int32)
    ind = C.c_int(-1)
    m = ref_lib().ref_minvect_int(_i(v), v.shape[0], C.byref(ind))
    return int(m), int(ind.value)


def ref_feature_matching(q, t):
    """left_nn / right_nn / mutual pairs: the loops of src/lidar_odometry.cpp:212-242 around the reference's minVect"""
    q = np.ascontiguousarray(q, dtype=np.uint64).reshape(-1, 6)
    t = np.ascontiguousarray(t, dtype=np.uint64).reshape(-1, 6)
    left = np.empty(q.shape[0], np.int32)
    right = np.empty(t.shape[0], np.int32)
    pairs = np.empty((q.shape[0], 2), np.int32)
    n = ref_lib().ref_feature_matching(_u(q), q.shape[0], _u(t), t.shape[0], _i(left), _i(right), _i(pairs))
    return dict(left_idx=left, right_idx=right, pairs=pairs[:n].copy())


def ref_preprocess(azimuth_deg, vertical_deg, distance, ring_deg, vert_init=-0.6, lowpt_th=-1950.0, select=None, save_selected=True):
    """myslam::Preprocessor::run of the reference (src/preprocess.cpp:213-223, compiled unchanged) on one rotation of
    returns; vert_init / lowpt_th as the SLAM driver sets them (test/odometry_test.cpp:118-119)"""
    az = np.ascontiguousarray(azimuth_deg, dtype=np.float64)
    ve = np.ascontiguousarray(vertical_deg, dtype=np.float64)
    di = np.ascontiguousarray(distance, dtype=np.uint16)
    ring = np.ascontiguousarray(ring_deg, dtype=np.float64)
    out = np.empty((max(az.size, 1), 3), np.float32)
    dp = C.POINTER(C.c_double)
    sel = None if select is None else np.ascontiguousarray(select, dtype=np.int32)
    n = ref_lib().ref_preprocess_select(az.ctypes.data_as(dp), ve.ctypes.data_as(dp), di.ctypes.data_as(C.POINTER(C.c_ushort)), az.size,
                                        ring.ctypes.data_as(dp), ring.size, vert_init, lowpt_th, None if sel is None else _i(sel),
                                        0 if sel is None else sel.size, 0 if sel is None else 1, 1 if save_selected else 0, _f(out), out.shape[0])
    return out[:n].copy()


class RefCb:
    """a `bshot cb` object of the reference living across frames (persistent cloud1_normals, include/bshot_bits.h:59)"""

    def __init__(self):
        self.h = ref_lib().ref_cb_create()

    def __del__(self):
        if getattr(self, "h", None):
            ref_lib().ref_cb_destroy(self.h)
            self.h = None

    def compute_descriptors(self, xyz, kp, radius=3000.0):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        kp = np.ascontiguousarray(kp, dtype=np.float32).reshape(-1, 3)
        k, n = kp.shape[0], xyz.shape[0]
        bits = np.empty((k, 6), np.uint64)
        shot = np.empty((k, 352), np.float32)
        rf = np.empty((k, 9), np.float32)
        normals = np.empty((n, 4), np.float32)
        ref_lib().ref_cb_compute_descriptors(self.h, _f(xyz), n, _f(kp), k, radius, _u(bits), _f(shot), _f(rf), _f(normals))
        return dict(bits=bits, shot=shot, rf=rf, normals=normals)

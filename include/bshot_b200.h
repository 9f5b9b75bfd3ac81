/*
 * bshot_b200.h -- C ABI of the B200-native (sm_100a) B-SHOT feature front end.
 *
 * Drop-in boundary for the per-frame front end of TingKaiChen/B-SHOT-SLAM.  The reference has no
 * FFI layer: control enters the path through C++ member calls on `class bshot`
 * (include/bshot_bits.h:30-281) made by LidarOdometry::extractKeypoints / computeDescriptors /
 * featureMatching (src/lidar_odometry.cpp:51,173,186).  Every entry point below names the
 * reference code it replaces.  The reference-named host C++ shims (headers under b-shot-slam_b200/host:
 * `bshot`, `bshot_descriptor`, `minVect`, `Frame`, `Keypoint`, `Map`) forward to this ABI; see
 * INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; all functions return 0 on success, a negative BSHOT_E_* code
 *     otherwise and never throw.  bshot_last_error() returns a thread-local message.
 *   - units are millimetres (reference: src/preprocess.cpp:46).
 *   - a descriptor record is 48 bytes = std::bitset<352> of libstdc++ (6 x u64 little endian, bit i
 *     in word i/64 at position i%64, bits 352..383 zero) == bshot_descriptor
 *     (include/bshot_bits.h:23-27), so std::vector<bshot_descriptor>::data() can be passed as is.
 *   - a context owns its device buffers (sized at creation, no per-frame allocation) and ONE
 *     CUDA stream.  It is not thread safe.  Host-buffer calls are synchronous on return.
 *     *_dev / *_resident calls take device pointers / use resident data and are asynchronous on
 *     the context stream.
 *   - there is NO CPU fallback: every call fails with BSHOT_E_CUDA when no sm_100 device is
 *     usable.
 */
#ifndef BSHOT_B200_H
#define BSHOT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSHOT_B200_VERSION 100

enum {
    BSHOT_OK = 0,
    BSHOT_E_INVALID = -1,   /* bad argument */
    BSHOT_E_CAPACITY = -2,  /* input larger than the context was created for */
    BSHOT_E_CUDA = -3,      /* CUDA runtime / launch failure, or no usable device */
    BSHOT_E_STATE = -4      /* call order violated (e.g. descriptors before a cloud) */
};

/* seg-ratio variants, LidarOdometry::sr_type_ "CV"/"CVS"/"CVSN" (src/lidar_odometry.cpp:83,98,109) */
enum { BSHOT_SR_CV = 0, BSHOT_SR_CVS = 1, BSHOT_SR_CVSN = 2 };

/* normals placement.  REFERENCE reproduces include/bshot_bits.h:58-59,79-81: the K keypoint
 * normals are stored at indices 0..K-1 of an N-sized, zero-initialised, persistent array that
 * SHOT then indexes by SURFACE point.  FULL computes a normal for every surface point (what the
 * commented-out lines include/bshot_bits.h:89-90 did). */
enum { BSHOT_NORMALS_REFERENCE = 0, BSHOT_NORMALS_FULL = 1 };

typedef struct bshot_ctx bshot_ctx;

/* one frame's parameters; bshot_params_default() fills the reference's literals */
typedef struct bshot_params {
    float kp_radius;     /* 3000  src/lidar_odometry.cpp:68 */
    int kp_max_nn;       /* 300   src/lidar_odometry.cpp:70 */
    int sr_type;         /* CV    src/lidar_odometry.cpp:6  */
    int top_k;           /* 600   src/lidar_odometry.cpp:138-142 */
    float normal_radius; /* 3000  src/lidar_odometry.cpp:174 */
    int normal_max_nn;   /* 300   include/bshot_bits.h:68 */
    int normals_mode;    /* BSHOT_NORMALS_REFERENCE */
    float shot_radius;   /* 3000  src/lidar_odometry.cpp:175 */
} bshot_params;

void bshot_params_default(bshot_params* p);

int bshot_version(void);
const char* bshot_last_error(void);

/* ---- context ------------------------------------------------------------------------------ */
/* replaces the `bshot cb` member of LidarOdometry (include/lidar_odometry.h:57) */
int bshot_ctx_create(bshot_ctx** out, int device, size_t max_points, size_t max_keypoints,
                     size_t max_targets);
void bshot_ctx_destroy(bshot_ctx* ctx);
/* the context's cudaStream_t (for CUDA-event timing by the caller) */
void* bshot_ctx_stream(bshot_ctx* ctx);
int bshot_ctx_sync(bshot_ctx* ctx);
/* forget all cross-frame state, as if `cb` were freshly constructed: the persistent normals array
 * (include/bshot_bits.h:59 resize keeps old entries), the previous frame's descriptors and the
 * map shard.  The cloud, keypoints and capacity stay. */
int bshot_ctx_reset(bshot_ctx* ctx);

/* ---- a1: cloud upload + voxel-hash build -------------------------------------------------- */
/* replaces LidarOdometry::setSrcFrame (src/lidar_odometry.cpp:29-41) + `cb.cloud1 = src_pcl_`
 * (:159) and the three pcl::KdTreeFLANN builds (:53-54, include/bshot_bits.h:52-53, PCL SHOT).
 * stride_bytes is 12 (Eigen::Vector3f) or 16 (pcl::PointXYZ). */
int bshot_set_cloud(bshot_ctx* ctx, const float* xyz, size_t n, size_t stride_bytes);

/* ---- a2+a3: seg-ratio keypoint detector ---------------------------------------------------- */
/* replaces LidarOdometry::extractKeypoints loop + sort + top-K (src/lidar_odometry.cpp:61-153).
 * Keypoints come out in ascending-ratio order (ties: descending point index), the top_k highest.
 * idx_out / ratio_out / kp_xyz_out (3 floats each) may be NULL; *count_out <= top_k.
 * The selected keypoints become the context's current keypoints (cb.cloud1_keypoints, :161). */
int bshot_detect_keypoints(bshot_ctx* ctx, float radius, int max_nn, int sr_type, int top_k,
                           int* idx_out, float* ratio_out, float* kp_xyz_out, int* count_out);
/* all N seg-ratios (NaN where the reference skips the point, :63,:121) */
int bshot_seg_ratio(bshot_ctx* ctx, float radius, int max_nn, int sr_type, float* ratio_out);
/* explicit keypoints instead of the detector: `cb.cloud1_keypoints = ...` (:161) */
int bshot_set_keypoints(bshot_ctx* ctx, const float* kp_xyz, size_t k, size_t stride_bytes);

/* ---- a4: normals ---------------------------------------------------------------------------- */
/* replaces bshot::calculate_normals (include/bshot_bits.h:43-94).  normals_out (may be NULL):
 * N x 4 floats (nx,ny,nz,curvature) indexed by surface point, i.e. cloud1_normals. */
int bshot_compute_normals(bshot_ctx* ctx, int mode, float radius, int max_nn, float* normals_out);
/* normals of arbitrary query points (Q x 4), no placement; the per-query body of :61-88 */
int bshot_query_normals(bshot_ctx* ctx, const float* q_xyz, size_t nq, float radius, int max_nn,
                        float* normals_out);
/* upload cloud1_normals verbatim (N x 4 floats) */
int bshot_set_normals(bshot_ctx* ctx, const float* normals4, size_t n);

/* ---- a5+a6+a7: SHOT LRF, SHOT352 histogram, B-SHOT bits ------------------------------------ */
/* replaces bshot::calculate_SHOT (include/bshot_bits.h:113-135) and bshot::compute_bshot (:138-142)
 * for the current keypoints and normals.  Any output may be NULL:
 *   bits_out  K x 6 u64 (bshot_descriptor)       shot_out  K x 352 f32 (pcl::SHOT352::descriptor)
 *   rf_out    K x 9 f32 (pcl::SHOT352::rf)        nn_out    K ints, neighbours within radius
 * sum_nn_out = sum of nn_out (the algorithmic-bytes unit of SURVEY 8d). */
int bshot_compute_shot(bshot_ctx* ctx, float radius, uint64_t* bits_out, float* shot_out,
                       float* rf_out, int* nn_out, long long* sum_nn_out);
/* LRF only (PCL SHOTLocalReferenceFrameEstimation behind include/bshot_bits.h:117-128) */
int bshot_compute_lrf(bshot_ctx* ctx, float radius, float* rf_out, int* valid_nn_out);
/* replaces bshot::compute_bshot_from_SHOT (include/bshot_bits.h:144-278) on caller floats.
 * shot: k records of 352 floats, stride_floats apart (361 for pcl::SHOT352). */
int bshot_binarize(bshot_ctx* ctx, const float* shot, size_t k, size_t stride_floats,
                   uint64_t* bits_out);
/* replaces LidarOdometry::computeDescriptors (src/lidar_odometry.cpp:173-184): normals + SHOT +
 * B-SHOT for the current keypoints in one call. */
int bshot_compute_descriptors(bshot_ctx* ctx, const bshot_params* p, uint64_t* bits_out);

/* ---- a10+a11: Hamming correspondence search ------------------------------------------------ */
/* replaces the two brute-force loops + minVect (src/lidar_odometry.cpp:212-232,
 * include/bshot_bits.h:6-20): left_idx[i] = argmin_k popcount(q_i ^ t_k) (first minimum wins),
 * right_idx[k] = argmin_i popcount(t_k ^ q_i).  Also returns the runner-up (second by
 * (distance, index)).  Any output may be NULL; right_idx == NULL skips the T x Q pass. */
int bshot_match(bshot_ctx* ctx, const uint64_t* q, size_t nq, const uint64_t* t, size_t nt,
                int* left_idx, int* left_dist, int* left_idx2, int* left_dist2, int* right_idx);
/* replaces the mutual-NN filter (src/lidar_odometry.cpp:234-242) fused with the search: writes
 * (index_query, index_match) pairs in ascending index_query; *count_out <= nq. Only the targets
 * that are some query's nearest neighbour get their reverse search (Q x Q instead of T x Q). */
int bshot_match_mutual(bshot_ctx* ctx, const uint64_t* q, size_t nq, const uint64_t* t, size_t nt,
                       int* pairs_out, int* dist_out, int* count_out);

/* ---- whole frame (the reference's extractKeypoints -> computeDescriptors -> featureMatching) */
/* Runs voxel build, detector, normals, SHOT, B-SHOT on `xyz` and matches the new descriptors
 * against the previous frame's descriptors kept in the context (first frame: against itself,
 * src/lidar_odometry.cpp:187-194), i.e. test/odometry_test.cpp:174-180 up to the RANSAC call.
 * One H2D copy in, one D2H copy out.  Outputs may be NULL.  kp_idx_out / bits_out hold up to
 * p->top_k records, pairs_out up to top_k (query,match) pairs. */
int bshot_process_frame(bshot_ctx* ctx, const bshot_params* p, const float* xyz, size_t n,
                        size_t stride_bytes, int* kp_idx_out, uint64_t* bits_out, int* n_kp_out,
                        int* pairs_out, int* n_pairs_out);
/* same work on a cloud that is already in device memory (d_xyz: device pointer); asynchronous on
 * the context stream, no host copies -- the HBM-resident timing leg of bench.py. */
int bshot_process_frame_dev(bshot_ctx* ctx, const bshot_params* p, const void* d_xyz, size_t n,
                            size_t stride_bytes);
/* download the results of the last bshot_process_frame_dev (synchronous) */
int bshot_fetch_frame(bshot_ctx* ctx, int top_k, int* kp_idx_out, uint64_t* bits_out, int* n_kp_out,
                      int* pairs_out, int* n_pairs_out);

/* ---- GPU-resident global map + frame-to-map flow (RUN status of featureMatching) ------------------------------------
 * Replaces, on the device: Keypoint::createKeypoint's 10 mm snap (src/keypoint.cpp:23-32), Map::addKeypoint
 * (src/mymap.cpp:4-26: 10 m blocks, reject when a stored keypoint of the block within 800 mm is at least as salient, else
 * insert / overwrite at the same position), LidarOdometry::updateMap (src/lidar_odometry.cpp:343-358: every keypoint of the
 * frame in order, transformed by the pose), Map::getKeypoints (src/mymap.cpp:28-74: blocks of the +-range cube, x outermost,
 * z innermost) and the target assembly of featureMatching (src/lidar_odometry.cpp:197-206: map subset, then the reference
 * frame's keypoints transformed by its pose).  Inside a block keypoints come out in insertion order (the reference's
 * unordered_map order is implementation defined).  pose3x4 = row-major [R|T], NULL = identity.
 * Per-frame call order of a host that keeps RANSAC / ICP (src/lidar_odometry.cpp:251-301) on the CPU:
 *   bshot_extract_frame -> bshot_match_frame_to_map -> (host: pose) -> bshot_gmap_update_from_frame(pose) -> bshot_frame_commit */
int bshot_gmap_create(bshot_ctx* ctx, size_t max_entries, size_t max_blocks);
int bshot_gmap_reset(bshot_ctx* ctx);
/* entries stored so far; dropped = keypoints lost to a full table / pool (0 unless the capacities are too small) */
int bshot_gmap_size(bshot_ctx* ctx, size_t* entries_out, size_t* dropped_out);
/* Map::addKeypoint for n keypoints given on the host (xyz n x 3 floats, in keypoint order), n <= max_keypoints per call */
int bshot_gmap_add(bshot_ctx* ctx, const float* xyz, const float* seg_ratio, const uint64_t* desc, size_t n, const float* pose3x4);
/* updateMap for the frame extracted last (keypoints, seg-ratios and descriptors are still on the device); asynchronous */
int bshot_gmap_update_from_frame(bshot_ctx* ctx, const float* pose3x4);
/* Map::getKeypoints to host buffers (cap records); *n_out = keypoints in range */
int bshot_gmap_get_keypoints(bshot_ctx* ctx, const float pos[3], float range, float* xyz_out, uint64_t* desc_out, size_t cap, size_t* n_out);
/* extractKeypoints + computeDescriptors of one scan (no matching); outputs hold up to p->top_k records, any may be NULL */
int bshot_extract_frame(bshot_ctx* ctx, const bshot_params* p, const float* xyz, size_t n, size_t stride_bytes, int* kp_idx_out,
                        float* kp_xyz_out, float* seg_ratio_out, uint64_t* bits_out, int* n_kp_out);
/* featureMatching in RUN status for the frame extracted last: targets = map keypoints within `range` of ref_pos, then the
 * committed (previous) frame transformed by ref_pose3x4; mutual nearest neighbours -> pairs_out (index_query, index_match
 * into that target set).  target_xyz_out (optional, target_cap x 3 floats) = positions of the targets for the host's RANSAC. */
int bshot_match_frame_to_map(bshot_ctx* ctx, const float ref_pos[3], float range, const float* ref_pose3x4, int* pairs_out,
                             int* n_pairs_out, size_t* n_targets_out, float* target_xyz_out, size_t target_cap);
/* the frame extracted last becomes the reference frame (passSrc2Ref, src/lidar_odometry.cpp:43-47); asynchronous */
int bshot_frame_commit(bshot_ctx* ctx);

/* ---- RANSAC correspondence rejection (the consumer of the path's output) ----------------------------------------------------
 * Replaces Ransac_based_Rejection.setMaximumIterations(2000) / setInputSource / setInputTarget / setInlierThreshold(1500) /
 * setInputCorrespondences / getCorrespondences (src/lidar_odometry.cpp:251-261), i.e. PCL 1.8's
 * CorrespondenceRejectorSampleConsensus: RandomSampleConsensus over SampleConsensusModelRegistration with PCL's deterministic
 * sample sequence (mt19937 seeded 12345), Umeyama on three pairs, adaptive stop, first best wins.  All max_iterations + 1
 * hypotheses are scored in parallel on the device; the adaptive loop is replayed over the scores.
 * pairs: n_pairs x (index_query into src, index_match into tgt).  Outputs: the surviving correspondences in their original
 * order (capacity n_pairs), their number, the best transformation (row-major 4x4, float) and the iterations PCL would have run.
 * Fewer than 3 correspondences / inliers: everything is kept and the transformation is the identity (PCL's behaviour). */
int bshot_ransac(bshot_ctx* ctx, const float* src_xyz, size_t n_src, const float* tgt_xyz, size_t n_tgt, const int* pairs,
                 size_t n_pairs, int max_iterations, float inlier_threshold, int* inlier_pairs_out, int* n_inliers_out,
                 float* transform4x4_out, int* iterations_out);

/* ---- ICP refinement and the estimation gate ------------------------------- */
/* Replaces pcl::IterativeClosestPoint<PointXYZ, PointXYZ> with PCL's defaults as LidarOdometry::evaluateEstimation uses it
 * (src/lidar_odometry.cpp:283-291): the n_src source points are first moved by pre4x4 (row-major 4x4, NULL = identity; the
 * reference's transformPointCloud with T_est), then aligned to the n_tgt target points: nearest neighbour correspondences,
 * Umeyama without scaling, at most max_iterations (PCL default 10) rounds.  final4x4_out = icp.getFinalTransformation();
 * state_out: 0 not converged, 1 iteration limit, 2 transformation epsilon, 3 absolute MSE, 5 fewer than 3 correspondences
 * (pcl::registration::DefaultConvergenceCriteria::ConvergenceState); mse_out = mean squared correspondence distance of the
 * last round.  Any output pointer may be NULL. */
int bshot_icp(bshot_ctx* ctx, const float* src_xyz, size_t n_src, const float* tgt_xyz, size_t n_tgt, const float* pre4x4,
              int max_iterations, float* final4x4_out, int* iterations_out, int* state_out, double* mse_out);
/* LidarOdometry::evaluateEstimation (src/lidar_odometry.cpp:267-296) without its printing: T_ij = T_ref^-1 * T_ransac, heading
 * change acos(T_ij(1,1)) and translation |t_ij|; the estimate is rejected (T_est = T_ref, *should_update_map_out = 0) when the
 * heading moved more than 10 degrees, the translation more than 1200 mm or fewer than 15 correspondences survived RANSAC;
 * otherwise T_est = T_ransac.  run_icp != 0: T_best = ICP(src_kp moved by T_est -> tgt_kp) * T_est; else T_best = T_ransac. */
int bshot_evaluate_estimation(bshot_ctx* ctx, const float* T_ransac4x4, const float* T_ref4x4, int n_correspondences,
                              const float* src_kp_xyz, size_t n_src, const float* tgt_kp_xyz, size_t n_tgt, int run_icp,
                              float* T_best4x4_out, int* should_update_map_out, float* h_diff_rad_out, float* t_diff_out,
                              int* icp_iterations_out);

/* ---- scan preprocessor ---------------------------------------------------- */
/* Replaces myslam::Preprocessor::setLasers + setVerticalAngles + run + getPointCloud (include/preprocess.h:26-33,
 * src/preprocess.cpp:213-223): one rotation of raw returns -> ground, self-car and occluded returns removed -> the cloud
 * the front end consumes, in the reference's order (azimuth-major, vertical angle ascending).
 * azimuth_deg / vertical_deg / distance: the n returns' velodyne::Laser fields (VelodyneCapture.h: azimuth and vertical in
 * degrees as double, distance in 2 mm units), sorted by azimuth as `capture.retrieve(lasers, true)` delivers them.
 * ring_deg: the sensor's nv vertical angles (setVerticalAngles).  vert_init_rad / lowpt_th: the constructor arguments
 * (the SLAM driver uses -0.6 rad and -1950 mm, test/odometry_test.cpp:32-33).  xyz_out: cap x 3 floats (mm); *n_out = number of kept points (if it
 * exceeds cap only the first cap are written and BSHOT_E_CAPACITY is returned). */
int bshot_preprocess(bshot_ctx* ctx, const double* azimuth_deg, const double* vertical_deg, const unsigned short* distance,
                     size_t n, const double* ring_deg, size_t nv, double vert_init_rad, double lowpt_th, float* xyz_out,
                     size_t cap, size_t* n_out);

/* the same with the preprocessor's point selection (setSelectedPoints / haveSelectList / saveSelectPoints,
 * include/preprocess.h:27-29): select_list holds indices into the returns; with have_select_list != 0 a return is
 * "selected" when the list names it, otherwise every return is; only returns whose selection equals save_selected != 0
 * are written (the reference's defaults: no list, save_selected = true). */
int bshot_preprocess_select(bshot_ctx* ctx, const double* azimuth_deg, const double* vertical_deg, const unsigned short* distance,
                            size_t n, const double* ring_deg, size_t nv, double vert_init_rad, double lowpt_th,
                            const int* select_list, size_t n_select, int have_select_list, int save_selected, float* xyz_out,
                            size_t cap, size_t* n_out);

/* Preprocessor::run + extractKeypoints + computeDescriptors of one rotation (test/odometry_test.cpp:143-177) in one call:
 * the preprocessed cloud stays on the device and feeds the front end directly.  cloud_xyz_out (cloud_cap x 3 floats) may be
 * NULL when the host does not need the cloud; the keypoint outputs are those of bshot_extract_frame (kp_idx indexes the
 * preprocessed cloud).  Follow with bshot_match_frame_to_map / bshot_ransac / bshot_evaluate_estimation /
 * bshot_gmap_update_from_frame / bshot_frame_commit as after bshot_extract_frame. */
int bshot_extract_scan(bshot_ctx* ctx, const bshot_params* p, const double* azimuth_deg, const double* vertical_deg,
                       const unsigned short* distance, size_t n, const double* ring_deg, size_t nv, double vert_init_rad,
                       double lowpt_th, float* cloud_xyz_out, size_t cloud_cap, size_t* n_points_out, int* kp_idx_out,
                       float* kp_xyz_out, float* seg_ratio_out, uint64_t* bits_out, int* n_kp_out);

/* ---- sharded map matching (north_star multi-GPU piece) -------------------------------------- */
/* The accumulated map descriptors (Map::getKeypoints output, include/mymap.h:34-38) are split
 * across ranks; each rank keeps its shard resident.  global_base = index of the shard's first
 * record in the global target array. */
int bshot_map_reset(bshot_ctx* ctx);
int bshot_map_append(bshot_ctx* ctx, const uint64_t* desc, size_t n);
/* same with descriptors that already live on the device (48-byte records) */
int bshot_map_append_dev(bshot_ctx* ctx, const void* d_desc, size_t n);
int bshot_map_size(bshot_ctx* ctx, size_t* n_out);
/* candidate record per query produced by one shard: packed keys (distance << 32 | global index),
 * 0xFFFFFFFFFFFFFFFF = none; rq = best query for target k1 among this call's queries. */
typedef struct bshot_cand {
    uint64_t k1;
    uint64_t k2;
    uint32_t rq;
    uint32_t pad;
} bshot_cand;
/* device pointers, asynchronous on the context stream: d_q (nq x 48 B) against the resident shard
 * -> d_cand_out (nq records).  with_rq != 0 also fills rq (Q x Q reverse pass). */
int bshot_match_shard_dev(bshot_ctx* ctx, const void* d_q, size_t nq, uint64_t global_base,
                          int with_rq, void* d_cand_out);
/* generic device-pointer matcher against caller-owned targets */
int bshot_match_dev(bshot_ctx* ctx, const void* d_q, size_t nq, const void* d_t, size_t nt,
                    uint64_t global_base, int with_rq, void* d_cand_out);
/* merge `nranks` candidate arrays (rank-major, nranks x nq records, e.g. an all-gather result)
 * by (distance, global index); d_out: nq merged records. mutual iff merged rq == query index. */
int bshot_merge_cands_dev(bshot_ctx* ctx, const void* d_cands, size_t nranks, size_t nq,
                          void* d_out);
/* Post-merge reverse pass, sharded: for the queries whose MERGED winner (d_merged, nq records) lies in this
 * rank's shard, find the best query of that target; d_rq_out (nq x u32) gets the query index there and
 * 0xFFFFFFFF elsewhere.  An all-reduce(MIN) of d_rq_out over the ranks followed by bshot_apply_rq_dev
 * completes the records with Q*Q/ranks pairs of work per rank (vs Q*Q for with_rq = 1 above). */
int bshot_reverse_owned_dev(bshot_ctx* ctx, const void* d_q, size_t nq, uint64_t global_base,
                            const void* d_merged, void* d_rq_out);
int bshot_apply_rq_dev(bshot_ctx* ctx, void* d_cands, const void* d_rq, size_t nq);
/* Peer-memory variant of the exchange (symmetric buffers mapped on every rank, e.g. torch symmetric memory over
 * NVLink / NVSwitch): instead of an all-gather, every rank STORES its nq records into slot `rank` of every rank's
 * gather buffer (d_peer_ptrs = device array of nranks pointers to buffers of nranks x nq records); instead of an
 * all-reduce, the owner of a winner stores rq[query] into every rank's rq array (d_peer_rq_ptrs = device array of
 * nranks pointers to nq x uint32).  The caller places one cross-rank barrier after each of the two calls. */
int bshot_push_cands_dev(bshot_ctx* ctx, const void* d_cands, size_t nq, const void* d_peer_ptrs, int nranks, int rank);
/* the barrier: d_peer_flag_ptrs = device array of nranks pointers to each rank's flag array (>= nranks uint32, zero
 * before the first barrier).  Every rank must issue the same sequence of barriers.  A rank that never arrives makes the
 * others give up after ~2 s instead of hanging; bshot_peer_barrier_timeouts then reports the epoch (0 = none). */
int bshot_peer_barrier_dev(bshot_ctx* ctx, const void* d_peer_flag_ptrs, int nranks, int rank);
/* start a new flag array: every rank calls this (collectively) right after its zeroed flag array is allocated, so the
 * barrier epoch always matches the flags -- a second matcher on the same context, or one created after a failed call,
 * starts in step.  Also clears the timeout record. */
int bshot_peer_barrier_reset(bshot_ctx* ctx);
int bshot_peer_barrier_timeouts(bshot_ctx* ctx, unsigned* epoch_out);
int bshot_reverse_owned_push_dev(bshot_ctx* ctx, const void* d_q, size_t nq, uint64_t global_base,
                                 const void* d_merged, const void* d_peer_rq_ptrs, int nranks, int rank);
/* ---- multi-rank exchange behind the C ABI (one process per GPU on one node; no Python / torch / NCCL needed) ------
 * Replaces what featureMatching (src/lidar_odometry.cpp:197-232) does against Map::getKeypoints' output when the
 * map is sharded over GPUs: every rank owns a contiguous range of the global target array (bshot_map_append) and
 * calls bshot_match_map_sharded[_dev] with the same queries; every rank receives all nq complete records (top-2 over
 * the WHOLE map with the reference's first-minimum tie-break, rq = best query of the winner).
 * Set-up, once:  bshot_comm_create on every rank -> bshot_comm_export -> exchange the 64-byte handles by any means
 * (pipe, file, MPI, torch.distributed) -> bshot_comm_import.  Ranks that live in ONE process pass device pointers
 * instead (bshot_comm_region / bshot_comm_import_ptrs; the caller enables peer access between the devices).
 * Per call (six kernel launches, stream-ordered, no host synchronisation): shard search; the merged top-2 records are
 * STORED into every rank's gather buffer over NVLink and a flag is released by the last CTA; the consumer kernel
 * acquires the flags, merges the ranks' records and selects the winners that live in its own shard; their best query
 * (Q x Q / ranks pairs) is stored into every rank's rq array behind a second flag set.  A rank that never arrives
 * makes the others give up after ~2 s: the next bshot_comm_check / bshot_match_map_sharded returns BSHOT_E_STATE. */
typedef struct bshot_ipc_handle { unsigned char bytes[64]; } bshot_ipc_handle;
int bshot_comm_create(bshot_ctx* ctx, int rank, int nranks, size_t max_queries);
int bshot_comm_export(bshot_ctx* ctx, bshot_ipc_handle* handle_out);
int bshot_comm_import(bshot_ctx* ctx, const bshot_ipc_handle* handles /* nranks entries, own entry ignored */);
int bshot_comm_region(bshot_ctx* ctx, void** d_region_out, size_t* bytes_out);
int bshot_comm_import_ptrs(bshot_ctx* ctx, void* const* d_regions /* nranks device pointers, own entry ignored */);
int bshot_comm_destroy(bshot_ctx* ctx);
int bshot_comm_check(bshot_ctx* ctx);
/* device pointers, asynchronous on the context stream: d_q (nq x 48 B) -> d_cand_out (nq bshot_cand records) */
int bshot_match_map_sharded_dev(bshot_ctx* ctx, const void* d_q, size_t nq, uint64_t global_base, void* d_cand_out);
/* host buffers, synchronous; checks the arrival of every rank */
int bshot_match_map_sharded(bshot_ctx* ctx, const uint64_t* q, size_t nq, uint64_t global_base, bshot_cand* cand_out);

/* host-buffer convenience over the three calls above for a single rank */
int bshot_match_map(bshot_ctx* ctx, const uint64_t* q, size_t nq, uint64_t global_base,
                    bshot_cand* cand_out);

/* ---- instrumentation ------------------------------------------------------------------------ */
/* number of kernels this library launched on the context since creation (bench.py gpu_launches) */
unsigned long long bshot_launch_count(bshot_ctx* ctx);
/* per-stage device times of whole-frame calls.  When enabled, CUDA events are recorded on the
 * context stream around each stage; bshot_stage_times synchronises and returns the LAST frame's
 * milliseconds: [0] voxel build, [1] seg-ratio, [2] top-K, [3] normals, [4] SHOT+B-SHOT,
 * [5] matching, [6] whole frame, [7] unused. */
int bshot_ctx_enable_timing(bshot_ctx* ctx, int on);
int bshot_stage_times(bshot_ctx* ctx, float ms_out[8]);
/* work counters of the last frame: [0] sum over points of min(#neighbours, max_nn) in the detector,
 * [1] same for the normals queries, [2] sum of SHOT neighbour counts, [3] keypoints */
int bshot_frame_counters(bshot_ctx* ctx, unsigned long long out[4]);
/* raw device counters of the last frame, for tuning the block-tiled neighbourhood kernel: [0],[1] as above,
 * [2] tiles staged, [3] tile points swept (sum over query attempts of the tile size), [4] query attempts,
 * [5] attempts whose sphere held fewer than max_nn points (retried with a larger tile), [6] queries handed to the
 * warp-per-query fallback, [7] query blocks processed (all tiled launches of the frame together) */
int bshot_debug_counters(bshot_ctx* ctx, unsigned long long out[8]);
/* Which kernel computes the Hamming distance matrix of every search on this context (all of them return the same
 * bit-exact (distance, lowest index) winners as minVect, include/bshot_bits.h:6-20):
 *   BSHOT_MATCHER_AUTO (default)  by problem size: tensor-core pipeline for searches of >= 2^21 pairs, XOR + POPC below
 *   BSHOT_MATCHER_POPC            XOR + POPC top-2 kernel, query tile in registers, targets through TMA (hamming.cu)
 *   BSHOT_MATCHER_TC              tcgen05 kind::i8, one CTA-serial tile at a time (hamming_tc.cu)
 *   BSHOT_MATCHER_TC_PIPELINED    warp-specialised tcgen05 pipeline, query tiles in tensor memory (hamming_tc2.cu)
 *   BSHOT_MATCHER_TC_PIPELINED_SMEM  the same with the query tiles in shared memory
 * The environment variable BSHOT_MATCH_TC=<kind> sets the initial value at bshot_ctx_create. */
#define BSHOT_MATCHER_AUTO (-1)
#define BSHOT_MATCHER_POPC 0
#define BSHOT_MATCHER_TC 1
#define BSHOT_MATCHER_TC_PIPELINED 2
#define BSHOT_MATCHER_TC_PIPELINED_SMEM 3
int bshot_set_matcher(bshot_ctx* ctx, int kind);
/* POPC-pipe microbenchmark: returns measured POPC32 instructions/s over the whole GPU */
int bshot_popc_peak(bshot_ctx* ctx, double* popc_per_s_out);

#ifdef __cplusplus
}
#endif
#endif

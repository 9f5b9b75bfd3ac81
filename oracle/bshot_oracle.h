/*
 * bshot_oracle.h -- C interface of the CPU ORACLE for the B-SHOT front end.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product path
 * (b-shot-slam_b200/csrc, include/bshot_b200.h) never links, loads or calls anything in oracle/.
 *
 * PARITY PIN (what is and is not anchored to the reference):
 *   - PINNED to reference-compiled code: the reference's own header include/bshot_bits.h is compiled
 *     UNCHANGED into oracle/_ref/libbshot_ref.so (oracle/ref_shim.cpp + oracle/pcl_stub, recipe in
 *     oracle/Makefile).  compute_bshot_from_SHOT (:144-278), minVect (:6-20), the 48-byte
 *     std::bitset<352> record and the matching loops built on them are checked against it live and through
 *     the committed vectors tests/golden/ref_pin.npz (tests/test_ref_pin.py); calculate_normals (:43-94)
 *     and calculate_SHOT (:113-135) run from the reference header too, which pins their control flow,
 *     NaN branches and the keypoint-ordinal / persistent-buffer placement quirk.
 *   - UNPINNED: the PCL arithmetic behind those calls.  The reference ships no golden vectors, and PCL
 *     (>= 1.7.2, unvendored) / Eigen / FLANN are not installable here, so kd-tree radius search,
 *     computeCentroid, computePointNormal/eigen33, SHOTLocalReferenceFrameEstimation and SHOTEstimation
 *     are restated from the published PCL 1.8 algorithms and pinned only by
 *     truth tables, analytic known answers, numpy.linalg.eigh / scipy cKDTree cross-checks and an
 *     independent numpy restatement (tests/test_oracle_*.py).  src/lidar_odometry.cpp (detector, top-K)
 *     needs Sophus/g2o/OpenCV and cannot be compiled: restated exactly from :61-153,186-242.
 */
#ifndef BSHOT_ORACLE_H
#define BSHOT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_cloud orc_cloud;

enum { ORC_SR_CV = 0, ORC_SR_CVS = 1, ORC_SR_CVSN = 2 };
enum { ORC_TIE_STDSORT = 0, ORC_TIE_DETERMINISTIC = 1 };

/* surface cloud + uniform grid (replaces the three pcl::KdTreeFLANN builds,
 * src/lidar_odometry.cpp:53-54, include/bshot_bits.h:52-53 and PCL-internal in SHOT) */
orc_cloud* orc_cloud_create(const float* xyz, size_t n, size_t stride_floats);
void orc_cloud_destroy(orc_cloud* c);
size_t orc_cloud_size(const orc_cloud* c);

/* pcl::KdTreeFLANN::radiusSearch semantics: all points with fp32
 * squared distance < radius^2, ascending by (sqd, index); max_nn > 0 keeps the max_nn nearest.
 * Returns the count (<= cap written). */
int orc_radius_search(const orc_cloud* c, const float q[3], float radius, int max_nn,
                      int* out_idx, float* out_sqd, int cap);

/* src/lidar_odometry.cpp:61-126 : seg-ratio per point (NaN where the reference skips). */
void orc_seg_ratio(const orc_cloud* c, float radius, int max_nn, int sr_type, float* ratio_out,
                   int threads);

/* src/lidar_odometry.cpp:131-153 : sort ascending, keep last top_k. Returns count. */
int orc_select_keypoints(const float* ratio, size_t n, int top_k, int tie_mode, int* idx_out,
                         float* ratio_out);

/* include/bshot_bits.h:61-88 loop body: normal (nx,ny,nz,curvature) per query point. */
void orc_normals(const orc_cloud* c, const float* q_xyz, size_t nq, float radius, int max_nn,
                 float* normal4_out, int threads);

/* pcl::computePointNormal(cloud, indices, n, curvature) alone (Appendix A.3): normal + curvature of the
 * points idx[0..n_idx) of xyz, accumulated in index order, NOT flipped; NaN when n_idx < 3.
 * (what the PCL stand-ins of oracle/pcl_stub call when the reference header runs, oracle/ref_shim.cpp) */
void orc_point_normal_indices(const float* xyz, size_t n, size_t stride_floats, const int* idx, int n_idx,
                              float out4[4]);

/* PCL SHOTLocalReferenceFrameEstimation::getLocalRF (Appendix A.4). rf = [x;y;z] rows. */
void orc_lrf(const orc_cloud* c, const float* kp_xyz, size_t nk, float radius, float* rf9_out,
             int* valid_nn_out, int threads);

/* PCL SHOTEstimation::computePointSHOT (Appendix A.5).  normals4 is indexed by SURFACE index
 * (size n*4). rf9_in may be NULL (then the LRF is computed). Returns sum of neighbour counts. */
long long orc_shot(const orc_cloud* c, const float* kp_xyz, size_t nk, float radius,
                   const float* normals4, const float* rf9_in, float* shot352_out,
                   float* rf9_out, int* nn_out, int threads);

/* include/bshot_bits.h:144-278 */
void orc_bshot(const float* shot352, size_t nk, uint64_t* bits6_out);

/* src/lidar_odometry.cpp:212-232 + minVect include/bshot_bits.h:6-20 (first minimum wins).
 * Any output pointer may be NULL. top-2: second best by (distance, index) order. */
void orc_match(const uint64_t* q, size_t nq, const uint64_t* t, size_t nt, int* left_idx,
               int* left_dist, int* left_idx2, int* left_dist2, int* right_idx, int threads);

/* src/lidar_odometry.cpp:234-242 ; writes (index_query,index_match) pairs, returns count */
int orc_mutual(const int* left_idx, size_t nq, const int* right_idx, int* pairs_out);

/* whole computeDescriptors (src/lidar_odometry.cpp:173-184) in the reference's REFERENCE
 * normals mode (quirk: keypoint normals stored at indices 0..K-1 of an N-sized zero array,
 * include/bshot_bits.h:58-59,79-81) or FULL mode (normals for every surface point).
 * mode: 0 = REFERENCE, 1 = FULL.  Outputs may be NULL. Returns sum of SHOT neighbour counts. */
long long orc_compute_descriptors(const orc_cloud* c, const float* kp_xyz, size_t nk,
                                  float radius, int max_nn, int mode, uint64_t* bits6_out,
                                  float* shot352_out, float* rf9_out, float* normals4_out,
                                  int threads);

/* global keypoint map: Keypoint::createKeypoint (src/keypoint.cpp:23-32), Map::addKeypoint / getKeypoints / getBlockID
 * (src/mymap.cpp:4-26, 28-74, 103-112), updateMap's pose (src/lidar_odometry.cpp:351; pose12 = row-major [R|T] or NULL).
 * Keypoints of one block come out in insertion order (the reference's unordered_map order is implementation defined). */
typedef struct orc_map orc_map;
orc_map* orc_map_create(void);
void orc_map_destroy(orc_map* m);
size_t orc_map_size(const orc_map* m);
void orc_map_add(orc_map* m, const float* xyz, const float* ratio, const uint64_t* desc, size_t n, const float* pose12);
size_t orc_map_get(orc_map* m, const float pos[3], float range, float* xyz_out, uint64_t* desc_out, size_t cap);

/* RANSAC correspondence rejection, src/lidar_odometry.cpp:251-261 = PCL 1.8 CorrespondenceRejectorSampleConsensus
 * (RandomSampleConsensus + SampleConsensusModelRegistration; mt19937(12345) sample sequence, Umeyama on three pairs,
 * adaptive stop).  pairs: n_pairs x (index_query, index_match).  Returns the number of surviving correspondences
 * (written to inlier_pairs_out in original order); UNPINNED restatement of PCL (not installable here). */
int orc_ransac(const float* src_xyz, const float* tgt_xyz, const int* pairs, size_t n_pairs, int max_iterations,
               double threshold, int* inlier_pairs_out, float transform_out[16], int* iterations_out);

/* src/lidar_odometry.cpp:283-291: pcl::IterativeClosestPoint (PCL 1.8 defaults) on src moved by pre4x4 (NULL = identity);
 * returns the convergence state (1 iteration limit, 2 transformation, 3 absolute mse, 5 no correspondences).
 * UNPINNED restatement of PCL. */
int orc_icp(const float* src_xyz, size_t n_src, const float* tgt_xyz, size_t n_tgt, const float* pre4x4, int max_iterations,
            float final_out[16], int* iterations_out, double* mse_out);
/* LidarOdometry::evaluateEstimation (src/lidar_odometry.cpp:267-296); returns shouldUpdateMap */
int orc_evaluate_estimation(const float* T_j, const float* T_i, int n_corr, const float* src_kp, size_t n_src, const float* tgt_kp,
                            size_t n_tgt, int run_icp, float T_best[16], float* h_diff_out, float* t_diff_out);

/* symmetric 3x3 eigen decomposition used by orc_lrf (double, ascending), exposed for tests */
void orc_eigh3(const double m[9], double evals[3], double evecs_cols[9]);
/* pcl::eigen33 smallest-eigenpair (fp32), exposed for tests */
void orc_eigen33_smallest(const float m[9], float* eval, float evec[3]);

#ifdef __cplusplus
}
#endif
#endif

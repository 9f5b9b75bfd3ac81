// stages.h -- internal stage entry points (one per SURVEY 8a row); all asynchronous on ctx->stream.
#pragma once
#include "common.cuh"

namespace bshot {

// a1: caller layout in d_raw (stride 3 or 4 floats) -> d_pts, voxel grid, d_sorted   (grid.cu)
int grid_build(Ctx* c, const float* d_raw, size_t n, int stride_floats);

// a2/a3: seg-ratio for every point -> d_ratio, d_keys ; top-K -> d_kp_idx/d_kp_ratio/d_kp/d_kp_count (detect.cu)
int detect_seg_ratio(Ctx* c, float radius, int max_nn, int sr_type, int fuse = 0, int gate_top_k = 0);  // fuse: 1 FULL normals, 2 covariance sums for the keypoint normals
int detect_topk(Ctx* c, int top_k);

// block-tiled exact neighbourhoods of cloud points (tilek.cu): seg-ratio scores and / or normals; d_flags = nullptr: every
// point, else per cell-sorted position the output slot (>= 0) of the points that want a normal
int tile_neighbourhoods(Ctx* c, int sr_type, bool seg, int nrm, float radius, int max_nn, const int* d_flags, float4* d_nrm_out, int gate_top_k = 0);
bool tile_path_ok(const Ctx* c, int max_nn);
// normals of the detector's keypoints at their ordinals, from the radii the detector kept (tilek.cu)
int tile_keypoint_normals(Ctx* c, float radius, int max_nn, float4* d_nrm_out);
// normals of the queries the tiled kernel put on the fallback list (normals.cu)
int normals_fallback_list(Ctx* c, float radius, int max_nn, const int* d_flags, float4* d_out);

// a4: normals (normals.cu). normals_query: q (float4 xyz_) -> out (nx,ny,nz,curvature); in-place allowed
int normals_query(Ctx* c, const float4* d_q, size_t nq, float radius, int max_nn, float4* d_out);
int normals_compute(Ctx* c, int mode, float radius, int max_nn);

// a5/a6/a7: LRF + SHOT352 + B-SHOT for the current keypoints (shot.cu)
int shot_compute(Ctx* c, float radius, bool lrf_only, bool write_shot);
int binarize(Ctx* c, const float* d_shot, size_t k, uint64_t* d_bits);

// a10/a11 (hamming.cu)
int hamming_top2(Ctx* c, const void* d_q, size_t nq, const void* d_t, size_t nt, unsigned long long global_base,
                 bshot_cand* d_out, unsigned* d_colmin = nullptr, const unsigned* d_nq = nullptr, const unsigned* d_nt = nullptr);
int hamming_reverse_owned(Ctx* c, const void* d_q, size_t nq, const void* d_t, size_t nt, unsigned long long global_base,
                          const bshot_cand* d_merged, unsigned* d_rq_out, const void* d_peer_rq = nullptr, unsigned nranks = 1,
                          unsigned rank = 0);
int hamming_peer_barrier(Ctx* c, const void* d_peer_flags, unsigned nranks, unsigned rank);
int hamming_push_cands(Ctx* c, const bshot_cand* d_cands, size_t nq, const void* d_peer_ptrs, unsigned nranks, unsigned rank);
int hamming_apply_rq(Ctx* c, bshot_cand* d_cand, const unsigned* d_rq, size_t nq);
int hamming_match_rq(Ctx* c, const void* d_q, size_t nq, const void* d_t, size_t nt, unsigned long long global_base,
                     bshot_cand* d_out, const unsigned* d_nq = nullptr, const unsigned* d_nt = nullptr);
int hamming_reverse(Ctx* c, const void* d_q, size_t nq, const void* d_t, unsigned long long global_base,
                    bshot_cand* d_cand, const unsigned* d_nq = nullptr);
int hamming_merge_cands(Ctx* c, const void* d_cands, size_t nranks, size_t nq, void* d_out);
int hamming_unpack(Ctx* c, const bshot_cand* d_cand, size_t nq, int* idx1, int* d1, int* idx2, int* d2);
int hamming_mutual_pairs(Ctx* c, const bshot_cand* d_cand, size_t nq, int* d_pairs3, int* d_count, const unsigned* d_nq = nullptr);
int popc_peak(Ctx* c, double* out);
// sharded map match over peer memory (six launches, hamming.cu)
size_t comm_region_bytes(const Comm& m);
int comm_set_peers(Ctx* c, void* const* region_ptrs);
int hamming_preload_sharded();
int hamming_match_sharded(Ctx* c, const void* d_q, size_t nq, unsigned long long global_base, bshot_cand* d_out);

// tensor-core distance matrix (hamming_tc.cu): per-split top-2 partials in c->d_partial
int hamming_tc_partials(Ctx* c, const void* d_q, size_t nq, const void* d_t, size_t nt, unsigned long long global_base, const unsigned* d_nq,
                        const unsigned* d_nt, unsigned* nsplit_out);
int hamming_tc2_preload();
int hamming_tc2_partials(Ctx* c, const void* d_q, size_t nq, const void* d_t, size_t nt, unsigned long long global_base, const unsigned* d_nq,
                         const unsigned* d_nt, unsigned* nsplit_out, size_t live_q = 0);

// GPU-resident global map (gmap.cu)
int gmap_create(Ctx* c, size_t max_entries, size_t max_blocks);
void gmap_free(Ctx* c);
int gmap_reset(Ctx* c);
int gmap_update(Ctx* c, const float4* d_kp, const float* d_ratio, const uint64_t* d_bits, const int* d_count, size_t n_cap, const float* pose12);
int gmap_gather(Ctx* c, const float pos[3], float range, const float4* d_ref_kp, const uint64_t* d_ref_bits, const int* d_ref_count,
                size_t ref_cap, const float* ref_pose12, uint64_t* d_t_out, size_t out_cap, unsigned* d_total);

// RANSAC correspondence rejection (ransac.cu)
int ransac_run(Ctx* c, const float* src_xyz, const float* tgt_xyz, const int* pairs, size_t n_pairs, int max_iterations, double threshold,
               int* inlier_pairs_out, int* n_inliers_out, float* transform_out, int* iterations_out);

// ICP refinement (icp.cu)
int icp_run(Ctx* c, const float* src_xyz, size_t n_src, const float* tgt_xyz, size_t n_tgt, const float* pre4x4, int max_iterations,
            float* final4x4_out, int* iterations_out, int* state_out, double* mse_out);

// scan preprocessor (preprocess.cu); scratch_reserve grows one of the context's on-demand scratch buffers
int scratch_reserve(Ctx* c, int which, size_t bytes);
int preprocess_run(Ctx* c, const double* az_deg, const double* vert_deg, const unsigned short* dist, size_t n, const double* ring_deg, size_t nv,
                   double vert_init, double lowpt_th, const unsigned char* sel, int save_sel, float* xyz_out, size_t cap, size_t* n_out,
                   const float** d_xyz_out = nullptr);

// whole frame on the resident cloud (frame.cu)
int frame_extract(Ctx* c, const bshot_params* p, const float* d_raw, size_t n, int stride_floats);
int frame_commit(Ctx* c, size_t k);
int frame_run(Ctx* c, const bshot_params* p, const float* d_raw, size_t n, int stride_floats);

}  // namespace bshot

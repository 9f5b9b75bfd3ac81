// frame.cu -- the reference's per-frame call order on the resident cloud, fully asynchronous:
// extractKeypoints -> computeDescriptors -> featureMatching (test/odometry_test.cpp:174-194,
// src/lidar_odometry.cpp:51-265 up to the RANSAC call).
#include "stages.h"

namespace bshot {

__global__ void copy_prev_kernel(const uint64_t* __restrict__ bits, const float4* __restrict__ kp, const int* __restrict__ count, unsigned cap,
                                 uint64_t* __restrict__ prev, float4* __restrict__ prev_kp, int* __restrict__ prev_count) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = *count;
    if (i < cap * 6 && (int)(i / 6) < k) prev[i] = bits[i];
    if (i < cap && (int)i < k) prev_kp[i] = kp[i];
    if (i == 0) *prev_count = k;
}

// the current frame becomes the reference frame (descriptors + keypoint positions + count stay on the device)
int frame_commit(Ctx* c, size_t k) {
    copy_prev_kernel<<<(unsigned)((k * 6 + 255) / 256), 256, 0, c->stream>>>(c->d_bits, c->d_kp, c->d_kp_count, (unsigned)k, c->d_prev_bits, c->d_prev_kp,
                                                                            c->d_prev_count);
    count_launch(c);
    c->n_prev = k;
    return check_launch("copy_prev_kernel");
}

// extractKeypoints + computeDescriptors (src/lidar_odometry.cpp:51-184) on a cloud in device memory, asynchronous
int frame_extract(Ctx* c, const bshot_params* p, const float* d_raw, size_t n, int stride_floats) {
    auto mark = [&](int i) {
        if (c->timing) cudaEventRecord(c->ev[i], c->stream);
    };
    BSHOT_CUDA_TRY(cudaMemsetAsync(c->d_counters, 0, 8 * sizeof(unsigned long long), c->stream));
    mark(0);
    BSHOT_TRY(grid_build(c, d_raw, n, stride_floats));
    mark(1);
    // FULL-mode normals use the detector's own neighbourhoods when the search parameters agree: one pass for both
    const bool same = p->normal_radius == p->kp_radius && p->normal_max_nn == p->kp_max_nn;
    const int fuse = !same ? 0 : (p->normals_mode == BSHOT_NORMALS_FULL ? 1 : 2);
    // sums for the deferred keypoint normals only where the previous frame's K-th score makes a keypoint likely
    const int gate = (c->gate_top_k == p->top_k && c->gate_sr == p->sr_type && c->gate_radius == p->kp_radius && c->gate_max_nn == p->kp_max_nn) ? p->top_k : 0;
    BSHOT_TRY(detect_seg_ratio(c, p->kp_radius, p->kp_max_nn, p->sr_type, fuse, gate));
    mark(2);
    BSHOT_TRY(detect_topk(c, p->top_k));
    c->gate_top_k = p->top_k; c->gate_sr = p->sr_type; c->gate_radius = p->kp_radius; c->gate_max_nn = p->kp_max_nn;
    mark(3);
    BSHOT_TRY(normals_compute(c, p->normals_mode, p->normal_radius, p->normal_max_nn));
    mark(4);
    BSHOT_TRY(shot_compute(c, p->shot_radius, false, false));
    mark(5);
    c->last_top_k = (size_t)p->top_k;
    return BSHOT_OK;
}

int frame_run(Ctx* c, const bshot_params* p, const float* d_raw, size_t n, int stride_floats) {
    auto mark = [&](int i) {
        if (c->timing) cudaEventRecord(c->ev[i], c->stream);
    };
    BSHOT_TRY(frame_extract(c, p, d_raw, n, stride_floats));
    if (c->ev_desc) BSHOT_CUDA_TRY(cudaEventRecord(c->ev_desc, c->stream));   // descriptors complete: bshot_process_frame starts their copy here
    // featureMatching: the initial frame is matched against itself (src/lidar_odometry.cpp:187-194),
    // later frames against the previous frame's descriptors.  Host-side counts are upper bounds
    // (top_k); the kernels trim queries AND targets by the device-side keypoint counts, so a frame that
    // yields fewer than top_k keypoints (the reference's `< 600` branch, :144-151) never matches stale records.
    const size_t k = (size_t)p->top_k;
    const bool initial = (c->n_prev == 0);
    const uint64_t* tgt = initial ? c->d_bits : c->d_prev_bits;
    const size_t nt = initial ? k : c->n_prev;
    const unsigned* d_nq = reinterpret_cast<const unsigned*>(c->d_kp_count);
    const unsigned* d_nt = reinterpret_cast<const unsigned*>(initial ? c->d_kp_count : c->d_prev_count);
    BSHOT_TRY(hamming_match_rq(c, c->d_bits, k, tgt, nt, 0, c->d_cand, d_nq, d_nt));
    BSHOT_TRY(hamming_mutual_pairs(c, c->d_cand, k, c->d_pairs, c->d_pair_count, d_nq));
    BSHOT_TRY(frame_commit(c, k));
    mark(6);
    c->ev_valid = c->timing;
    return BSHOT_OK;
}

}  // namespace bshot

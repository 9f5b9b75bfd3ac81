// normals.cu -- surface normals.
//
// Replaces bshot::calculate_normals (include/bshot_bits.h:43-94): per query point the nearest <= 300
// neighbours inside the radius (knn.cuh), pcl::computePointNormal = single-pass mean/covariance,
// smallest eigenvector by the closed-form pcl::eigen33, curvature, flip towards the origin
// (SURVEY Appendix A.3).  One warp per query.  REFERENCE mode reproduces the reference's placement
// quirk (normal of keypoint ordinal i stored at index i of the N-sized, persistent, zero-initialised
// array that SHOT indexes by surface point); FULL mode computes a normal per surface point.
#include <type_traits>

#include "knn.cuh"
#include "normal_math.cuh"
#include "stages.h"

namespace bshot {

constexpr int NM_WARPS = 4;
constexpr int NM_THREADS = NM_WARPS * 32;

// Warp-per-query kernel: queries that are NOT cloud points (bshot_query_normals, bshot_set_keypoints), the fallback list
// of the tiled kernel and max_nn beyond the tiled path.  pcl::computeMeanAndCovarianceMatrix as the reference runs it: nine
// fp32 accumulators in neighbour order (knn.cuh exact replay; selections larger than KN_EXACT_CAP keep fp64 sums).
//   list == nullptr : query j = queries[j], result -> out[j]
//   list != nullptr : query = sorted[list[j]], result -> out[flags ? flags[list[j]] : surface index]
__global__ void __launch_bounds__(NM_THREADS, 4)
normals_kernel(const GridParams* __restrict__ gp, const unsigned* __restrict__ cell_start,
               const float4* __restrict__ sorted, const float4* __restrict__ pts, const float4* queries, const int* __restrict__ nq_dev, unsigned nq,
               float radius, int max_nn, float4* out, unsigned long long* __restrict__ counters, const unsigned* __restrict__ list,
               const unsigned* __restrict__ list_n, const int* __restrict__ flags) {
    __shared__ KnnExactSmem smem[NM_WARPS];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const GridParams g = *gp;
    KnnWarpSmem& sm = smem[wid].k;
    const unsigned n_items = list ? *list_n : (nq_dev ? min(nq, (unsigned)max(*nq_dev, 0)) : nq);
    for (unsigned j = blockIdx.x * NM_WARPS + wid; j < n_items; j += gridDim.x * NM_WARPS) {
        float4 q;
        unsigned oi = j;
        if (list) {
            const unsigned pos = list[j];
            q = __ldg(sorted + pos);
            oi = flags ? (unsigned)flags[pos] : __float_as_uint(q.w);
        } else {
            q = queries[j];
        }
        const float nanf_ = __int_as_float(0x7FC00000);
        const bool finite = isfinite(q.x) && isfinite(q.y) && isfinite(q.z);
        int n = 0;
        double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        float a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        bool exact = false;
        if (finite) {
            KnnResult res = knn_select(g, cell_start, sorted, pts, q, radius, max_nn, sm, lane, [&](const float4 p) {
                s[0] += (double)__fmul_rn(p.x, p.x); s[1] += (double)__fmul_rn(p.x, p.y); s[2] += (double)__fmul_rn(p.x, p.z);
                s[3] += (double)__fmul_rn(p.y, p.y); s[4] += (double)__fmul_rn(p.y, p.z); s[5] += (double)__fmul_rn(p.z, p.z);
                s[6] += (double)p.x; s[7] += (double)p.y; s[8] += (double)p.z;
            });
            n = res.count;
            if (knn_sorted_selected(g, cell_start, sorted, q, res, sm, smem[wid].skeys, lane)) {
                knn_replay_in_order(pts, smem[wid].skeys, res.count, lane, [&](float x, float y, float z) {
                    a[0] = __fadd_rn(a[0], __fmul_rn(x, x)); a[1] = __fadd_rn(a[1], __fmul_rn(x, y)); a[2] = __fadd_rn(a[2], __fmul_rn(x, z));
                    a[3] = __fadd_rn(a[3], __fmul_rn(y, y)); a[4] = __fadd_rn(a[4], __fmul_rn(y, z)); a[5] = __fadd_rn(a[5], __fmul_rn(z, z));
                    a[6] = __fadd_rn(a[6], x); a[7] = __fadd_rn(a[7], y); a[8] = __fadd_rn(a[8], z);
                });
                exact = true;  // every lane holds the same sums
            } else {
#pragma unroll
                for (int k = 0; k < 9; ++k) a[k] = (float)warp_sum(s[k]);
            }
        }
        (void)exact;
        if (lane == 0) {
            atomicAdd(&counters[1], (unsigned long long)n);
            float4 o = make_float4(nanf_, nanf_, nanf_, nanf_);
            if (finite) o = normal_from_sums(a, n, q.x, q.y, q.z);
            out[oi] = o;
        }
        __syncwarp();
    }
}

__global__ void place_normals_kernel(const float4* __restrict__ src, const int* __restrict__ count, unsigned cap,
                                     float4* __restrict__ dst) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap && (int)i < *count) dst[i] = src[i];
}

// REFERENCE-mode normals of the detector's keypoints from the covariance sums the detector kept for every point
// (detect_seg_ratio, fuse == 2): one eigen-solve per keypoint, result at the keypoint's ordinal (include/bshot_bits.h:79-81).
// A keypoint whose sums are missing (its query went to the warp fallback) is appended to the fallback list instead.
__global__ void __launch_bounds__(128)
normals_from_sums_kernel(const float* __restrict__ qsums, const float4* __restrict__ pts, const int* __restrict__ kp_idx, const int* __restrict__ kp_count,
                         const unsigned* __restrict__ sorted_pos, float4* __restrict__ out, unsigned* __restrict__ fb_list, unsigned* __restrict__ fb_len,
                         unsigned long long* __restrict__ counters) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    int cnt_n = 0;
    if ((int)i < *kp_count) {
        const int idx = kp_idx[i];
        const float* s = qsums + 10 * (size_t)idx;
        const float cnt = s[9];
        if (cnt >= 1.0f) {
            float a[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) a[k] = s[k];
            const float4 q = pts[idx];
            out[i] = normal_from_sums(a, (int)cnt, q.x, q.y, q.z);
            cnt_n = (int)cnt;
        } else {
            fb_list[atomicAdd(fb_len, 1u)] = sorted_pos[idx];
            atomicAdd(&counters[6], 1ull);
        }
    }
    cnt_n = warp_sum(cnt_n);
    if ((threadIdx.x & 31) == 0 && cnt_n) atomicAdd(&counters[1], (unsigned long long)cnt_n);
}

// FULL mode: points that are not in the voxel table (non-finite coordinates) have no neighbourhood -> NaN normal
__global__ void nan_unbinned_normals_kernel(const unsigned* __restrict__ cell_of, unsigned n, float4* __restrict__ normals) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const float nanf_ = __int_as_float(0x7FC00000);
    if (i < n && cell_of[i] == 0xFFFFFFFFu) normals[i] = make_float4(nanf_, nanf_, nanf_, nanf_);
}

static int normals_launch(Ctx* c, const float4* d_q, const int* nq_dev, size_t nq, float radius, int max_nn, float4* d_out) {
    const unsigned ctas = (unsigned)((nq + NM_WARPS - 1) / NM_WARPS);
    normals_kernel<<<ctas, NM_THREADS, 0, c->stream>>>(c->d_grid, c->d_cell_start, c->d_sorted, c->d_pts, d_q, nq_dev, (unsigned)nq, radius,
                                                       max_nn, d_out, c->d_counters, nullptr, nullptr, nullptr);
    count_launch(c);
    return check_launch("normals_kernel");
}

int normals_fallback_list(Ctx* c, float radius, int max_nn, const int* d_flags, float4* d_out) {
    normals_kernel<<<(unsigned)c->sm_count * 4u, NM_THREADS, 0, c->stream>>>(c->d_grid, c->d_cell_start, c->d_sorted, c->d_pts, nullptr, nullptr, 0u,
                                                                            radius, max_nn, d_out, c->d_counters, c->d_fb_list,
                                                                            c->d_nblocks + 1, d_flags);
    count_launch(c);
    return check_launch("normals_kernel (fallback list)");
}

int normals_query(Ctx* c, const float4* d_q, size_t nq, float radius, int max_nn, float4* d_out) {
    if (nq == 0) return BSHOT_OK;
    if (!(radius > 0.0f)) { set_error("bad radius"); return BSHOT_E_INVALID; }
    return normals_launch(c, d_q, nullptr, nq, radius, max_nn, d_out);
}

int normals_compute(Ctx* c, int mode, float radius, int max_nn) {
    if (!(radius > 0.0f)) { set_error("bad radius"); return BSHOT_E_INVALID; }
    const bool tiled = tile_path_ok(c, max_nn) && c->n_points > 0;
    if (mode == BSHOT_NORMALS_FULL) {
        if (c->fused_normals && c->fused_radius == radius && c->fused_max_nn == max_nn) {
            // the detector already produced them from the same neighbourhoods (detect_seg_ratio, fuse_normals)
        } else if (tiled) {
                BSHOT_TRY(tile_neighbourhoods(c, 0, false, true, radius, max_nn, nullptr, c->d_normals));
            BSHOT_TRY(normals_fallback_list(c, radius, max_nn, nullptr, c->d_normals));
        } else {
            BSHOT_TRY(normals_query(c, c->d_pts, c->n_points, radius, max_nn, c->d_normals));
        }
        if (c->n_points) {
            nan_unbinned_normals_kernel<<<(unsigned)((c->n_points + 255) / 256), 256, 0, c->stream>>>(c->d_cell_of, (unsigned)c->n_points, c->d_normals);
            count_launch(c);
            BSHOT_TRY(check_launch("nan_unbinned_normals_kernel"));
        }
        c->normals_valid = c->n_points;
    } else {
        c->fused_normals = false;  // keypoint normals overwrite the prefix of d_normals
        const size_t k = std::min(c->n_kp, c->n_points);  // keypoint ordinal idx lands at surface index idx
        if (k) {
            if (tiled && c->kp_from_detector && c->sel_valid && c->fused_sums && c->fused_radius == radius && c->fused_max_nn == max_nn) {
                // the detector kept the covariance sums of every point's neighbourhood: no second search
                BSHOT_CUDA_TRY(cudaMemsetAsync(c->d_nblocks + 1, 0, sizeof(unsigned), c->stream));
                normals_from_sums_kernel<<<(unsigned)((k + 127) / 128), 128, 0, c->stream>>>(c->d_qsums, c->d_pts, c->d_kp_idx, c->d_kp_count, c->d_sorted_pos,
                                                                                           c->d_normals, c->d_fb_list, c->d_nblocks + 1, c->d_counters);
                count_launch(c);
                BSHOT_TRY(check_launch("normals_from_sums_kernel"));
                BSHOT_TRY(normals_fallback_list(c, radius, max_nn, c->d_kp_flag, c->d_normals));
            } else if (tiled && c->kp_from_detector && c->sel_valid && c->sel_radius == radius && c->sel_max_nn == max_nn) {
                // keypoints are cloud points the detector just searched with the same parameters: one private tile per
                // keypoint with the radius kept for it, straight into d_normals[ordinal] (include/bshot_bits.h:79-81)
                BSHOT_TRY(tile_keypoint_normals(c, radius, max_nn, c->d_normals));
                BSHOT_TRY(normals_fallback_list(c, radius, max_nn, c->d_kp_flag, c->d_normals));
            } else if (tiled && c->kp_from_detector) {
                // other search parameters: keypoints flagged per cell-sorted position, shared tiles
                BSHOT_TRY(tile_neighbourhoods(c, 0, false, true, radius, max_nn, c->d_kp_flag, c->d_normals));
                BSHOT_TRY(normals_fallback_list(c, radius, max_nn, c->d_kp_flag, c->d_normals));
            } else {
                BSHOT_TRY(normals_launch(c, c->d_kp, c->d_kp_count, k, radius, max_nn, c->d_qnormals));
                place_normals_kernel<<<(unsigned)((k + 255) / 256), 256, 0, c->stream>>>(c->d_qnormals, c->d_kp_count, (unsigned)k,
                                                                                         c->d_normals);
                count_launch(c);
                BSHOT_TRY(check_launch("place_normals_kernel"));
            }
        }
        c->normals_valid = std::max(c->normals_valid, k);
    }
    c->have_normals = true;
    return BSHOT_OK;
}

}  // namespace bshot

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
#include "stages.h"

namespace bshot {

constexpr int NM_WARPS = 4;
constexpr int NM_THREADS = NM_WARPS * 32;

__device__ __forceinline__ void compute_roots2(float b, float c, float roots[3]) {
    roots[0] = 0.0f;
    float d = (float)((double)__fmul_rn(b, b) - 4.0 * (double)c);
    if (d < 0.0f) d = 0.0f;
    const float sd = sqrtf(d);
    roots[2] = __fmul_rn(0.5f, __fadd_rn(b, sd));
    roots[1] = __fmul_rn(0.5f, __fsub_rn(b, sd));
}

// pcl::computeRoots (fp32, trigonometric closed form); products evaluated left to right, no FMA
__device__ void compute_roots(const float m[9], float roots[3]) {
    const float m00 = m[0], m01 = m[1], m02 = m[2], m11 = m[4], m12 = m[5], m22 = m[8];
    float c0 = __fmul_rn(__fmul_rn(m00, m11), m22);
    c0 = __fadd_rn(c0, __fmul_rn(__fmul_rn(__fmul_rn(2.0f, m01), m02), m12));
    c0 = __fsub_rn(c0, __fmul_rn(__fmul_rn(m00, m12), m12));
    c0 = __fsub_rn(c0, __fmul_rn(__fmul_rn(m11, m02), m02));
    c0 = __fsub_rn(c0, __fmul_rn(__fmul_rn(m22, m01), m01));
    float c1 = __fsub_rn(__fmul_rn(m00, m11), __fmul_rn(m01, m01));
    c1 = __fadd_rn(c1, __fmul_rn(m00, m22));
    c1 = __fsub_rn(c1, __fmul_rn(m02, m02));
    c1 = __fadd_rn(c1, __fmul_rn(m11, m22));
    c1 = __fsub_rn(c1, __fmul_rn(m12, m12));
    const float c2 = __fadd_rn(__fadd_rn(m00, m11), m22);
    if (fabsf(c0) < 1.1920929e-07f) {
        compute_roots2(c2, c1, roots);
        return;
    }
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = sqrtf(3.0f);
    const float c2_over_3 = __fmul_rn(c2, s_inv3);
    float a_over_3 = __fmul_rn(__fsub_rn(c1, __fmul_rn(c2, c2_over_3)), s_inv3);
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    const float inner = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, c2_over_3), c2_over_3), c1);
    const float half_b = __fmul_rn(0.5f, __fadd_rn(c0, __fmul_rn(c2_over_3, inner)));
    float qv = __fadd_rn(__fmul_rn(half_b, half_b), __fmul_rn(__fmul_rn(a_over_3, a_over_3), a_over_3));
    if (qv > 0.0f) qv = 0.0f;
    const float rho = sqrtf(-a_over_3);
    const float theta = __fmul_rn(atan2f(sqrtf(-qv), half_b), s_inv3);
    const float cos_theta = cosf(theta);
    const float sin_theta = sinf(theta);
    roots[0] = __fadd_rn(c2_over_3, __fmul_rn(__fmul_rn(2.0f, rho), cos_theta));
    roots[1] = __fsub_rn(c2_over_3, __fmul_rn(rho, __fadd_rn(cos_theta, __fmul_rn(s_sqrt3, sin_theta))));
    roots[2] = __fsub_rn(c2_over_3, __fmul_rn(rho, __fsub_rn(cos_theta, __fmul_rn(s_sqrt3, sin_theta))));
    if (roots[0] >= roots[1]) { const float t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
    if (roots[1] >= roots[2]) {
        const float t = roots[1]; roots[1] = roots[2]; roots[2] = t;
        if (roots[0] >= roots[1]) { const float u = roots[0]; roots[0] = roots[1]; roots[1] = u; }
    }
    if (roots[0] <= 0.0f) compute_roots2(c2, c1, roots);
}

__device__ __forceinline__ void cross_rn(const float* a, const float* b, float* o) {
    o[0] = __fsub_rn(__fmul_rn(a[1], b[2]), __fmul_rn(a[2], b[1]));
    o[1] = __fsub_rn(__fmul_rn(a[2], b[0]), __fmul_rn(a[0], b[2]));
    o[2] = __fsub_rn(__fmul_rn(a[0], b[1]), __fmul_rn(a[1], b[0]));
}

// pcl::eigen33(mat, eigenvalue, eigenvector): smallest eigenpair
__device__ void eigen33_smallest(const float mat[9], float& eigenvalue, float evec[3]) {
    float scale = 0.0f;
#pragma unroll
    for (int k = 0; k < 9; ++k) scale = fmaxf(scale, fabsf(mat[k]));
    if (scale <= 1.17549435e-38f) scale = 1.0f;
    float s[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) s[k] = mat[k] / scale;
    float roots[3];
    compute_roots(s, roots);
    eigenvalue = __fmul_rn(roots[0], scale);
    s[0] = __fsub_rn(s[0], roots[0]); s[4] = __fsub_rn(s[4], roots[0]); s[8] = __fsub_rn(s[8], roots[0]);
    float v1[3], v2[3], v3[3];
    cross_rn(s, s + 3, v1);
    cross_rn(s, s + 6, v2);
    cross_rn(s + 3, s + 6, v3);
    const float l1 = dot3_rn(v1[0], v1[1], v1[2], v1[0], v1[1], v1[2]);
    const float l2 = dot3_rn(v2[0], v2[1], v2[2], v2[0], v2[1], v2[2]);
    const float l3 = dot3_rn(v3[0], v3[1], v3[2], v3[0], v3[1], v3[2]);
    const float* v; float l;
    if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; }
    else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
    else { v = v3; l = l3; }
    const float inv = sqrtf(l);
    evec[0] = v[0] / inv; evec[1] = v[1] / inv; evec[2] = v[2] / inv;
}

// the 9 single-pass sums of pcl::computeMeanAndCovarianceMatrix over the selected neighbourhood (lane-partial)
__device__ __forceinline__ void normal_sums(const GridParams& g, const unsigned* __restrict__ cell_start,
                                            const float4* __restrict__ sorted, const float4* __restrict__ pts, const float4& q,
                                            float radius, int max_nn, KnnWarpSmem& sm, unsigned lane, double s[9], int& n) {
    const KnnResult res = knn_select(g, cell_start, sorted, pts, q, radius, max_nn, sm, lane, [&](const float4 p) {
        // products rounded to fp32 like PCL's accumulator inputs, summed in fp64
        s[0] += (double)__fmul_rn(p.x, p.x); s[1] += (double)__fmul_rn(p.x, p.y); s[2] += (double)__fmul_rn(p.x, p.z);
        s[3] += (double)__fmul_rn(p.y, p.y); s[4] += (double)__fmul_rn(p.y, p.z); s[5] += (double)__fmul_rn(p.z, p.z);
        s[6] += (double)p.x; s[7] += (double)p.y; s[8] += (double)p.z;
    });
    n = res.count;
}

// CACHED: the queries are keypoints chosen by the detector (w = surface index) and the detector searched with
// the same (radius, max_nn): their neighbourhoods are re-collected from the kept (rho2, threshold key).
template <bool CACHED, bool EXACT>
__global__ void __launch_bounds__(NM_THREADS, EXACT ? 4 : 8)
normals_kernel(const GridParams* __restrict__ gp, const unsigned* __restrict__ cell_start,
               const float4* __restrict__ sorted, const float4* __restrict__ pts, const float4* queries, const int* __restrict__ nq_dev, unsigned nq,
               float radius, int max_nn, float4* out, unsigned long long* __restrict__ counters,
               const float* __restrict__ sel_rho2, const unsigned long long* __restrict__ sel_thr) {
    using Smem = typename std::conditional<EXACT, KnnExactSmem, KnnWarpSmem>::type;
    __shared__ Smem smem[NM_WARPS];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned j = blockIdx.x * NM_WARPS + wid;
    if (j >= nq) return;
    if (nq_dev && (int)j >= *nq_dev) return;
    const GridParams g = *gp;
    KnnWarpSmem& sm = *reinterpret_cast<KnnWarpSmem*>(&smem[wid]);
    float4 q = queries[j];
    const float nanf_ = __int_as_float(0x7FC00000);
    const bool finite = isfinite(q.x) && isfinite(q.y) && isfinite(q.z);
    int n = 0;
    double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (finite) {
        if (CACHED) {
            const unsigned si = __float_as_uint(q.w);
            n = knn_collect_cached(g, cell_start, sorted, q, sel_rho2[si], sel_thr[si], sm, lane, [&](const float4 p) {
                s[0] += (double)__fmul_rn(p.x, p.x); s[1] += (double)__fmul_rn(p.x, p.y); s[2] += (double)__fmul_rn(p.x, p.z);
                s[3] += (double)__fmul_rn(p.y, p.y); s[4] += (double)__fmul_rn(p.y, p.z); s[5] += (double)__fmul_rn(p.z, p.z);
                s[6] += (double)p.x; s[7] += (double)p.y; s[8] += (double)p.z;
            });
        } else if constexpr (EXACT) {
            // pcl::computeMeanAndCovarianceMatrix as the reference runs it: nine fp32 accumulators, neighbour order
            KnnResult res = knn_select(g, cell_start, sorted, pts, q, radius, max_nn, sm, lane, [&](const float4 p) {
                s[0] += (double)__fmul_rn(p.x, p.x); s[1] += (double)__fmul_rn(p.x, p.y); s[2] += (double)__fmul_rn(p.x, p.z);
                s[3] += (double)__fmul_rn(p.y, p.y); s[4] += (double)__fmul_rn(p.y, p.z); s[5] += (double)__fmul_rn(p.z, p.z);
                s[6] += (double)p.x; s[7] += (double)p.y; s[8] += (double)p.z;
            });
            n = res.count;
            if (knn_sorted_selected(g, cell_start, sorted, q, res, sm, smem[wid].skeys, lane)) {
                float a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                knn_replay_in_order(pts, smem[wid].skeys, res.count, lane, [&](float x, float y, float z) {
                    a[0] = __fadd_rn(a[0], __fmul_rn(x, x)); a[1] = __fadd_rn(a[1], __fmul_rn(x, y)); a[2] = __fadd_rn(a[2], __fmul_rn(x, z));
                    a[3] = __fadd_rn(a[3], __fmul_rn(y, y)); a[4] = __fadd_rn(a[4], __fmul_rn(y, z)); a[5] = __fadd_rn(a[5], __fmul_rn(z, z));
                    a[6] = __fadd_rn(a[6], x); a[7] = __fadd_rn(a[7], y); a[8] = __fadd_rn(a[8], z);
                });
                // every lane holds the same sums: hand them to the common tail as lane 0's contribution
#pragma unroll
                for (int k = 0; k < 9; ++k) s[k] = (lane == 0) ? (double)a[k] : 0.0;
            }
        } else {
            normal_sums(g, cell_start, sorted, pts, q, radius, max_nn, sm, lane, s, n);
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) s[k] = warp_sum(s[k]);
    }
    if (lane != 0) return;
    atomicAdd(&counters[1], (unsigned long long)n);
    float4 o = make_float4(nanf_, nanf_, nanf_, nanf_);
    if (finite && n >= 3) {  // n == 0: include/bshot_bits.h:67-74 ; n < 3: computePointNormal guard
        const float fn = (float)n;
        float a[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) a[k] = (float)s[k] / fn;
        float cov[9];
        cov[0] = __fsub_rn(a[0], __fmul_rn(a[6], a[6]));
        cov[1] = __fsub_rn(a[1], __fmul_rn(a[6], a[7]));
        cov[2] = __fsub_rn(a[2], __fmul_rn(a[6], a[8]));
        cov[4] = __fsub_rn(a[3], __fmul_rn(a[7], a[7]));
        cov[5] = __fsub_rn(a[4], __fmul_rn(a[7], a[8]));
        cov[8] = __fsub_rn(a[5], __fmul_rn(a[8], a[8]));
        cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
        float ev, nv[3];
        eigen33_smallest(cov, ev, nv);
        const float eig_sum = __fadd_rn(__fadd_rn(cov[0], cov[4]), cov[8]);
        o.w = (eig_sum != 0.0f) ? fabsf(ev / eig_sum) : 0.0f;
        const float vx = 0.0f - q.x, vy = 0.0f - q.y, vz = 0.0f - q.z;
        const float cos_theta = __fadd_rn(__fadd_rn(__fmul_rn(vx, nv[0]), __fmul_rn(vy, nv[1])), __fmul_rn(vz, nv[2]));
        if (cos_theta < 0.0f) { nv[0] = -nv[0]; nv[1] = -nv[1]; nv[2] = -nv[2]; }
        o.x = nv[0]; o.y = nv[1]; o.z = nv[2];
    }
    out[j] = o;
}

__global__ void place_normals_kernel(const float4* __restrict__ src, const int* __restrict__ count, unsigned cap,
                                     float4* __restrict__ dst) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap && (int)i < *count) dst[i] = src[i];
}

static int normals_launch(Ctx* c, const float4* d_q, const int* nq_dev, size_t nq, float radius, int max_nn, float4* d_out,
                          bool cached = false) {
    const unsigned ctas = (unsigned)((nq + NM_WARPS - 1) / NM_WARPS);
    if (c->exact_sums)  // exact-order fp32 sums (knn.cuh): always a fresh, sorted selection
        normals_kernel<false, true><<<ctas, NM_THREADS, 0, c->stream>>>(c->d_grid, c->d_cell_start, c->d_sorted, c->d_pts, d_q, nq_dev,
                                                                        (unsigned)nq, radius, max_nn, d_out, c->d_counters, nullptr, nullptr);
    else if (cached)
        normals_kernel<true, false><<<ctas, NM_THREADS, 0, c->stream>>>(c->d_grid, c->d_cell_start, c->d_sorted, c->d_pts, d_q, nq_dev,
                                                                        (unsigned)nq, radius, max_nn, d_out, c->d_counters, c->d_sel_rho2,
                                                                        c->d_sel_thr);
    else
        normals_kernel<false, false><<<ctas, NM_THREADS, 0, c->stream>>>(c->d_grid, c->d_cell_start, c->d_sorted, c->d_pts, d_q, nq_dev,
                                                                         (unsigned)nq, radius, max_nn, d_out, c->d_counters, nullptr, nullptr);
    count_launch(c);
    return check_launch("normals_kernel");
}

int normals_query(Ctx* c, const float4* d_q, size_t nq, float radius, int max_nn, float4* d_out) {
    if (nq == 0) return BSHOT_OK;
    if (!(radius > 0.0f)) { set_error("bad radius"); return BSHOT_E_INVALID; }
    return normals_launch(c, d_q, nullptr, nq, radius, max_nn, d_out);
}

int normals_compute(Ctx* c, int mode, float radius, int max_nn) {
    if (!(radius > 0.0f)) { set_error("bad radius"); return BSHOT_E_INVALID; }
    if (mode == BSHOT_NORMALS_FULL) {
        BSHOT_TRY(normals_query(c, c->d_pts, c->n_points, radius, max_nn, c->d_normals));
        c->normals_valid = c->n_points;
    } else {
        const size_t k = std::min(c->n_kp, c->n_points);  // keypoint ordinal idx lands at surface index idx
        if (k) {
            // keypoints that came out of the detector with the same search parameters: reuse its neighbourhoods
            const bool cached = c->sel_valid && c->kp_from_detector && radius == c->sel_radius && max_nn == c->sel_max_nn;
            BSHOT_TRY(normals_launch(c, c->d_kp, c->d_kp_count, k, radius, max_nn, c->d_qnormals, cached));
            place_normals_kernel<<<(unsigned)((k + 255) / 256), 256, 0, c->stream>>>(c->d_qnormals, c->d_kp_count, (unsigned)k,
                                                                                     c->d_normals);
            count_launch(c);
            BSHOT_TRY(check_launch("place_normals_kernel"));
        }
        c->normals_valid = std::max(c->normals_valid, k);
    }
    c->have_normals = true;
    return BSHOT_OK;
}

}  // namespace bshot

// detect.cu -- seg-ratio ("SR") keypoint detector and top-K selection.
//
// Replaces the per-point loop of LidarOdometry::extractKeypoints (src/lidar_odometry.cpp:61-126)
// and the sort / keep-last-K that follows (:131-153).  One warp per point (taken in voxel order so
// neighbouring warps share cache lines): nearest-<=max_nn-inside-R selection (knn.cuh), centroid,
// then the CV / CVS / CVSN score.  Top-K is a single-CTA MSB radix select over 64-bit keys
// (ratio bits << 32 | ~index) followed by an in-shared-memory bitonic sort, so keypoints come out
// in ascending ratio order like the reference's `SegRatio.end()-600 .. end()` slice, with a
// deterministic tie-break (lower point index wins) where std::sort's is unspecified.
#include "knn.cuh"
#include "stages.h"

namespace bshot {

constexpr int DT_WARPS = 4;
constexpr int DT_THREADS = DT_WARPS * 32;

__global__ void __launch_bounds__(DT_THREADS)
seg_ratio_kernel(const GridParams* __restrict__ gp, const unsigned* __restrict__ cell_start,
                 const float4* __restrict__ sorted, unsigned n_total, float radius, int max_nn, int sr_type,
                 float* __restrict__ ratio, unsigned long long* __restrict__ keys,
                 unsigned long long* __restrict__ counters) {
    __shared__ KnnWarpSmem smem[DT_WARPS];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned j = blockIdx.x * DT_WARPS + wid;
    const GridParams g = *gp;
    if (j >= g.npoints || j >= n_total) return;
    // the sorted array holds only the binned (finite) points: [0, cell_start[ncells])
    if (j >= __ldg(cell_start + g.ncells)) return;
    KnnWarpSmem& sm = smem[wid];
    const float4 q = __ldg(sorted + j);
    const unsigned qi = __float_as_uint(q.w);
    const float nanf_ = __int_as_float(0x7FC00000);
    if (q.x == 0.0f && q.y == 0.0f && q.z == 0.0f) {  // src/lidar_odometry.cpp:63
        if (lane == 0) { ratio[qi] = nanf_; keys[qi] = 0ull; }
        return;
    }
    RowRange rr;
    const KnnResult res = knn_select(g, cell_start, sorted, q, radius, max_nn, sm, lane, rr);
    const float rho = sqrtf(res.rho2) * 1.0001f;
    bool cached = !res.batched;
    // centroid (pcl::computeCentroid, :76): sum in fp64, rounded once
    double sx = 0, sy = 0, sz = 0;
    knn_for_each(g, cell_start, sorted, q, rho, rr, sm, lane, cached, [&](const float4 p) {
        const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
        if (knn_selected(res, sqd, p.w)) { sx += p.x; sy += p.y; sz += p.z; }
    });
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    const float fn = (float)res.count;
    const float ctx = (float)sx / fn, cty = (float)sy / fn, ctz = (float)sz / fn;
    const float vx = __fsub_rn(q.x, ctx), vy = __fsub_rn(q.y, cty), vz = __fsub_rn(q.z, ctz);  // :79
    float seg;
    if (sr_type == BSHOT_SR_CV) {  // :83-97
        int pos = 0, neg = 0;
        knn_for_each(g, cell_start, sorted, q, rho, rr, sm, lane, cached, [&](const float4 p) {
            const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
            if (!knn_selected(res, sqd, p.w)) return;
            const float d = dot3_rn(vx, vy, vz, __fsub_rn(p.x, q.x), __fsub_rn(p.y, q.y), __fsub_rn(p.z, q.z));
            if (d > 0.0f) ++pos;
            else if (d < 0.0f) ++neg;
        });
        pos = warp_sum(pos);
        neg = warp_sum(neg);
        const float fp = (float)pos, fq = (float)neg;
        seg = 1.0f - fminf(fp, fq) / fmaxf(fp, fq);  // 0/0 -> NaN like the reference
        if (pos == 0 && neg == 0) seg = nanf_;
    } else {  // CVS :98-108, CVSN :109-119
        const float ctn = sqrtf(dot3_rn(vx, vy, vz, vx, vy, vz));
        double sum = 0.0;
        knn_for_each(g, cell_start, sorted, q, rho, rr, sm, lane, cached, [&](const float4 p) {
            const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
            if (!knn_selected(res, sqd, p.w)) return;
            const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
            const float dn = sqrtf(dot3_rn(dx, dy, dz, dx, dy, dz));
            if (ctn == 0.0f || dn == 0.0f) return;
            const float d = dot3_rn(vx, vy, vz, dx, dy, dz);
            sum += (sr_type == BSHOT_SR_CVS) ? (double)d : (double)(d / __fmul_rn(ctn, dn));
        });
        sum = warp_sum(sum);
        seg = fabsf((float)sum) / fn;
    }
    if (lane == 0) {
        atomicAdd(&counters[0], (unsigned long long)res.count);
        ratio[qi] = seg;
        keys[qi] = isnan(seg) ? 0ull : (((unsigned long long)__float_as_uint(seg) << 32) | (unsigned)(~qi));
    }
}

__global__ void mark_unbinned_kernel(const unsigned* __restrict__ cell_of, unsigned n, float* __restrict__ ratio,
                                     unsigned long long* __restrict__ keys) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && cell_of[i] == 0xFFFFFFFFu) { ratio[i] = __int_as_float(0x7FC00000); keys[i] = 0ull; }
}

// ---- top-K: single CTA radix select + bitonic sort -------------------------------------------------
constexpr int TK_THREADS = 1024;

__global__ void __launch_bounds__(TK_THREADS)
topk_kernel(const unsigned long long* __restrict__ keys, unsigned n, int top_k, unsigned sort_cap,
            const float4* __restrict__ pts, int* __restrict__ kp_idx, float* __restrict__ kp_ratio,
            float4* __restrict__ kp, int* __restrict__ kp_count) {
    extern __shared__ unsigned long long sbuf[];  // sort_cap keys
    __shared__ unsigned hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned s_rank, s_valid, s_fill;
    const unsigned tid = threadIdx.x;
    // number of valid (non-zero) keys
    if (tid == 0) { s_valid = 0; s_fill = 0; }
    __syncthreads();
    unsigned cv = 0;
    for (unsigned i = tid; i < n; i += TK_THREADS) cv += keys[i] != 0ull;
    cv = (unsigned)warp_sum((int)cv);
    if ((tid & 31) == 0) atomicAdd(&s_valid, cv);
    __syncthreads();
    const unsigned k_eff = min((unsigned)top_k, s_valid);
    if (tid == 0) *kp_count = (int)k_eff;
    if (k_eff == 0) return;
    // K-th largest key via MSB-first 8-bit radix select (rank counted from the top)
    if (tid == 0) { s_prefix = 0ull; s_rank = k_eff - 1; }
    __syncthreads();
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (unsigned b = tid; b < 256; b += TK_THREADS) hist[b] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        const unsigned long long mask = (shift == 56) ? 0ull : (~0ull << (shift + 8));
        for (unsigned i = tid; i < n; i += TK_THREADS) {
            const unsigned long long key = keys[i];
            if ((key & mask) == prefix) atomicAdd(&hist[(unsigned)(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned rank = s_rank;
            int d = 255;
            for (; d > 0; --d) {
                if (rank < hist[d]) break;
                rank -= hist[d];
            }
            s_rank = rank;
            s_prefix = prefix | ((unsigned long long)d << shift);
        }
        __syncthreads();
    }
    const unsigned long long kth = s_prefix;  // keys are distinct: exactly k_eff keys are >= kth
    for (unsigned i = tid; i < n; i += TK_THREADS) {
        const unsigned long long key = keys[i];
        if (key >= kth && key != 0ull) {
            const unsigned slot = atomicAdd(&s_fill, 1u);
            if (slot < sort_cap) sbuf[slot] = key;
        }
    }
    __syncthreads();
    unsigned m = 1;
    while (m < k_eff) m <<= 1;
    for (unsigned i = k_eff + tid; i < m; i += TK_THREADS) sbuf[i] = ~0ull;  // pad high
    __syncthreads();
    for (unsigned size = 2; size <= m; size <<= 1) {
        for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
            for (unsigned t = tid; t < (m >> 1); t += TK_THREADS) {
                const unsigned lo = 2 * t - (t & (stride - 1));
                const unsigned hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = sbuf[lo], b = sbuf[hi];
                if ((a > b) == up) { sbuf[lo] = b; sbuf[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (unsigned i = tid; i < k_eff; i += TK_THREADS) {
        const unsigned long long key = sbuf[i];
        const unsigned idx = ~(unsigned)(key & 0xFFFFFFFFull);
        kp_idx[i] = (int)idx;
        kp_ratio[i] = __uint_as_float((unsigned)(key >> 32));
        float4 p = pts[idx];
        p.w = __uint_as_float(idx);
        kp[i] = p;
    }
}

int detect_seg_ratio(Ctx* c, float radius, int max_nn, int sr_type) {
    const unsigned n = (unsigned)c->n_points;
    if (n == 0) return BSHOT_OK;
    if (sr_type < 0 || sr_type > 2) { set_error("bad sr_type %d", sr_type); return BSHOT_E_INVALID; }
    if (!(radius > 0.0f)) { set_error("bad radius"); return BSHOT_E_INVALID; }
    mark_unbinned_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_cell_of, n, c->d_ratio, c->d_keys);
    seg_ratio_kernel<<<(n + DT_WARPS - 1) / DT_WARPS, DT_THREADS, 0, c->stream>>>(
        c->d_grid, c->d_cell_start, c->d_sorted, n, radius, max_nn, sr_type, c->d_ratio, c->d_keys, c->d_counters);
    count_launch(c, 2);
    return check_launch("seg_ratio_kernel");
}

int detect_topk(Ctx* c, int top_k) {
    const unsigned n = (unsigned)c->n_points;
    unsigned cap = 1;
    while (cap < (unsigned)top_k) cap <<= 1;
    const size_t smem = sizeof(unsigned long long) * cap;
    if (smem > 200 * 1024) { set_error("top_k %d too large for the single-CTA sorter", top_k); return BSHOT_E_CAPACITY; }
    static bool attr_set = false;
    if (!attr_set) {
        BSHOT_CUDA_TRY(cudaFuncSetAttribute(topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    topk_kernel<<<1, TK_THREADS, smem, c->stream>>>(c->d_keys, n, top_k, cap, c->d_pts, c->d_kp_idx, c->d_kp_ratio,
                                                   c->d_kp, c->d_kp_count);
    count_launch(c);
    c->n_kp = (size_t)top_k;  // upper bound until the host reads d_kp_count
    c->have_kp = true;
    return check_launch("topk_kernel");
}

}  // namespace bshot

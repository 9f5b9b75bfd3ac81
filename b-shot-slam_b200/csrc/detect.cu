// detect.cu -- seg-ratio ("SR") keypoint detector and top-K selection.
//
// Replaces the per-point loop of LidarOdometry::extractKeypoints (src/lidar_odometry.cpp:61-126)
// and the sort / keep-last-K that follows (:131-153).  The scores come from the block-tiled kernel (tilek.cu,
// tile.cuh); the warp-per-point kernel below (nearest-<=max_nn-inside-R selection of knn.cuh, centroid, CV / CVS /
// CVSN score) is its fallback for the queries whose tile does not fit and for max_nn beyond the tiled path.  Both
// replay the reference's fp32 running sums in neighbour order.  Top-K is a multi-CTA histogram select over 64-bit keys
// (ratio bits << 32 | ~index) followed by a rank-by-counting scatter, so keypoints come out
// in ascending ratio order like the reference's `SegRatio.end()-600 .. end()` slice, with a
// deterministic tie-break (lower point index wins) where std::sort's is unspecified.
#include <type_traits>

#include "knn.cuh"
#include "stages.h"

namespace bshot {

constexpr int DT_WARPS = 4;
constexpr int DT_THREADS = DT_WARPS * 32;

#ifndef BSHOT_DT_MINBLOCKS
#define BSHOT_DT_MINBLOCKS 8
#endif

// seg-ratio of one point, all 32 lanes call.  skeys != nullptr: exact-order fp32 centroid (knn.cuh)
template <int SR>
__device__ __forceinline__ void seg_ratio_point(const GridParams& g, const unsigned* __restrict__ cell_start,
                                                const float4* __restrict__ sorted, const float4* __restrict__ pts,
                                                const float4& q, float radius, int max_nn, KnnWarpSmem& sm,
                                                unsigned long long* skeys, unsigned lane, float& seg, int& count) {
    const float nanf_ = __int_as_float(0x7FC00000);
    // centroid (pcl::computeCentroid, :76): sum in fp64, rounded once
    double sx = 0, sy = 0, sz = 0;
    KnnResult res = knn_select(g, cell_start, sorted, pts, q, radius, max_nn, sm, lane,
                                     [&](const float4 p) { sx += p.x; sy += p.y; sz += p.z; });
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    const float fn = (float)res.count;
    float ctx = (float)sx / fn, cty = (float)sy / fn, ctz = (float)sz / fn;
    bool exact = false;
    if (skeys && knn_sorted_selected(g, cell_start, sorted, q, res, sm, skeys, lane)) {
        // the reference's own arithmetic: fp32 running sums in ascending-distance order
        float fx = 0.0f, fy = 0.0f, fz = 0.0f;
        knn_replay_in_order(pts, skeys, res.count, lane, [&](float x, float y, float z) {
            fx = __fadd_rn(fx, x); fy = __fadd_rn(fy, y); fz = __fadd_rn(fz, z);
        });
        ctx = fx / fn; cty = fy / fn; ctz = fz / fn;
        exact = true;
    }
    const float vx = __fsub_rn(q.x, ctx), vy = __fsub_rn(q.y, cty), vz = __fsub_rn(q.z, ctz);  // :79
    if (SR == BSHOT_SR_CV) {  // :83-97
        int pos = 0, neg = 0;
        knn_for_each(g, cell_start, sorted, q, res.it, sm, lane, [&](const float4 p) {
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
    } else if (exact) {  // CVS / CVSN with the reference's fp32 running sum in neighbour order
        const float ctn = sqrtf(dot3_rn(vx, vy, vz, vx, vy, vz));
        float sum = 0.0f;
        knn_replay_in_order(pts, skeys, res.count, lane, [&](float x, float y, float z) {
            const float dx = __fsub_rn(x, q.x), dy = __fsub_rn(y, q.y), dz = __fsub_rn(z, q.z);
            const float dn = sqrtf(dot3_rn(dx, dy, dz, dx, dy, dz));
            if (ctn == 0.0f || dn == 0.0f) return;
            const float d = dot3_rn(vx, vy, vz, dx, dy, dz);
            sum = __fadd_rn(sum, (SR == BSHOT_SR_CVS) ? d : d / __fmul_rn(ctn, dn));
        });
        seg = fabsf(sum) / fn;
    } else {  // CVS :98-108, CVSN :109-119
        const float ctn = sqrtf(dot3_rn(vx, vy, vz, vx, vy, vz));
        double sum = 0.0;
        knn_for_each(g, cell_start, sorted, q, res.it, sm, lane, [&](const float4 p) {
            const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
            if (!knn_selected(res, sqd, p.w)) return;
            const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
            const float dn = sqrtf(dot3_rn(dx, dy, dz, dx, dy, dz));
            if (ctn == 0.0f || dn == 0.0f) return;
            const float d = dot3_rn(vx, vy, vz, dx, dy, dz);
            sum += (SR == BSHOT_SR_CVS) ? (double)d : (double)(d / __fmul_rn(ctn, dn));
        });
        sum = warp_sum(sum);
        seg = fabsf((float)sum) / fn;
    }
    count = res.count;
}

// SR = score type (compile time: the CV kernel carries no CVS / CVSN code).  One warp per binned point, taken
// in voxel order -- or, with `list`, per listed position of the cell-sorted array (tiled-path fallback).
template <int SR>
__global__ void __launch_bounds__(DT_THREADS, 4)
seg_ratio_kernel(const GridParams* __restrict__ gp, const unsigned* __restrict__ cell_start,
                 const float4* __restrict__ sorted, const float4* __restrict__ pts, unsigned n_total, float radius, int max_nn,
                 float* __restrict__ ratio, unsigned long long* __restrict__ keys,
                 unsigned long long* __restrict__ counters, const unsigned* __restrict__ list, const unsigned* __restrict__ list_n) {
    __shared__ KnnExactSmem smem[DT_WARPS];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const GridParams g = *gp;
    KnnWarpSmem& sm = smem[wid].k;
    unsigned long long* skeys = smem[wid].skeys;
    const unsigned n_items = list ? *list_n : min(__ldg(cell_start + g.ncells), n_total);
    const float nanf_ = __int_as_float(0x7FC00000);
    for (unsigned j = blockIdx.x * DT_WARPS + wid; j < n_items; j += gridDim.x * DT_WARPS) {
        const float4 q = __ldg(sorted + (list ? list[j] : j));
        const unsigned qi = __float_as_uint(q.w);
        if (q.x == 0.0f && q.y == 0.0f && q.z == 0.0f) {  // src/lidar_odometry.cpp:63
            if (lane == 0) { ratio[qi] = nanf_; keys[qi] = 0ull; }
            continue;
        }
        float seg;
        int count;
        seg_ratio_point<SR>(g, cell_start, sorted, pts, q, radius, max_nn, sm, skeys, lane, seg, count);
        if (lane == 0) {
            atomicAdd(&counters[0], (unsigned long long)count);
            ratio[qi] = seg;
            keys[qi] = isnan(seg) ? 0ull : (((unsigned long long)__float_as_uint(seg) << 32) | (unsigned)(~qi));
        }
    }
}

// points without a score: not binned (non-finite) or at the origin (src/lidar_odometry.cpp:63)
__global__ void mark_unscored_kernel(const float4* __restrict__ pts, const unsigned* __restrict__ cell_of, unsigned n,
                                     float* __restrict__ ratio, unsigned long long* __restrict__ keys, float* __restrict__ rho_hint,
                                     float radius) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    rho_hint[i] = radius;  // the tiled kernel overwrites it with the radius of the point's neighbourhood
    const float4 p = pts[i];
    if (cell_of[i] == 0xFFFFFFFFu || (p.x == 0.0f && p.y == 0.0f && p.z == 0.0f)) { ratio[i] = __int_as_float(0x7FC00000); keys[i] = 0ull; }
}

// ---- top-K (a3): multi-CTA histogram select + rank scatter ------------------------------------------
// keys are (ratio bits << 32 | ~index), 0 = no score.  The K largest keys come out in ASCENDING key
// order (ascending ratio like the reference's `SegRatio.end()-600 .. end()` slice; equal ratios: lower
// point index last, i.e. it survives a cut first).
//   tk_hist_kernel    4096-bin histogram of a monotone bin function of the ratio, privatised per CTA in
//                     shared memory and flushed with one global atomic per non-empty bin; the last CTA to
//                     finish (atomic ticket) scans it from the top, finds the bin b* that holds the K-th
//                     largest key, publishes {k_eff, b*, need = k_eff - #keys above b*} and clears the table
//   tk_compact_kernel keys above b* -> `sure` list, keys in b* -> `tie` list (warp-aggregated appends)
//   tk_rank_kernel    rank of every sure key among the sure keys and of every tie key among the tie keys
//                     by counting (keys are distinct, so ranks are a permutation); the `need` largest
//                     tie keys and all sure keys are scattered to their final sorted position
constexpr int TK_THREADS = 256;
constexpr int TK_BINS = 4096;
constexpr int TK_TILE = 2048;
constexpr int TK_ITEMS = 16;   // items ranked per CTA pass (tk_rank_kernel)
enum { TKS_VALID = 0, TKS_TICKET = 1, TKS_BIN = 2, TKS_KEFF = 3, TKS_NSURE = 4, TKS_NTIE = 5, TKS_NEED = 6 };

// monotone (non-decreasing in the key) bin: ratios in [0,1] -- all CV scores -- are spread uniformly over
// bins 0..2048, larger ones (CVS / CVSN sums) by exponent and 4 mantissa bits over 2048..4095
__device__ __forceinline__ unsigned tk_bin(unsigned long long key) {
    const unsigned bits = (unsigned)(key >> 32);
    const float r = __uint_as_float(bits);
    if (!(r > 1.0f)) return (unsigned)(fmaxf(r, 0.0f) * 2048.0f);  // NaN never reaches here (key 0 is skipped)
    return min(4095u, 2048u + ((bits - 0x3F800000u) >> 19));
}

__global__ void __launch_bounds__(TK_THREADS)
tk_hist_kernel(const unsigned long long* __restrict__ keys, unsigned n, int top_k, unsigned* __restrict__ hist,
               unsigned* __restrict__ state, int* __restrict__ kp_count) {
    __shared__ unsigned s_hist[TK_BINS];
    __shared__ unsigned s_warp[TK_THREADS / 32];
    __shared__ unsigned s_last, s_found_bin, s_found_above;
    const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (unsigned b = tid; b < TK_BINS; b += TK_THREADS) s_hist[b] = 0u;
    __syncthreads();
    unsigned cv = 0;
    for (unsigned i0 = blockIdx.x * TK_THREADS * 4; i0 < n; i0 += gridDim.x * TK_THREADS * 4) {
        unsigned long long k[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned i = i0 + u * TK_THREADS + tid;
            k[u] = (i < n) ? __ldg(keys + i) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (k[u] != 0ull) { ++cv; atomicAdd(&s_hist[tk_bin(k[u])], 1u); }
    }
    cv = (unsigned)warp_sum((int)cv);
    if (lane == 0 && cv) atomicAdd(&state[TKS_VALID], cv);
    __syncthreads();
    for (unsigned b = tid; b < TK_BINS; b += TK_THREADS) {
        const unsigned h = s_hist[b];
        if (h) atomicAdd(&hist[b], h);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&state[TKS_TICKET], 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- last CTA: thread t owns the 16 bins [4096 - 16 (t + 1), 4096 - 16 t), scanned from the top
    constexpr int PER = TK_BINS / TK_THREADS;
    const unsigned valid = __ldcg(state + TKS_VALID);
    const unsigned k_eff = min((unsigned)max(top_k, 0), valid);
    const unsigned hi = TK_BINS - tid * PER;  // exclusive upper bin
    unsigned h[PER], s = 0;
    {
        const uint4* h4 = reinterpret_cast<const uint4*>(hist + (hi - PER));
#pragma unroll
        for (int v = 0; v < PER / 4; ++v) {
            const uint4 a = __ldcg(h4 + v);
            h[4 * v] = a.x; h[4 * v + 1] = a.y; h[4 * v + 2] = a.z; h[4 * v + 3] = a.w;
            s += a.x + a.y + a.z + a.w;
        }
    }
    unsigned inc = s;  // inclusive scan over threads = keys in this thread's bins and all higher bins
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += up;
    }
    if (lane == 31) s_warp[wid] = inc;
    if (tid == 0) { s_found_bin = 0xFFFFFFFFu; s_found_above = 0; }
    __syncthreads();
    unsigned above = inc - s;
    for (unsigned w = 0; w < wid; ++w) above += s_warp[w];
    if (k_eff > 0 && above < k_eff && above + s >= k_eff) {  // exactly one thread: the K-th largest key is in its bins
        unsigned run = above;
#pragma unroll
        for (int b = PER - 1; b >= 0; --b) {
            if (run < k_eff && run + h[b] >= k_eff) { s_found_bin = hi - PER + (unsigned)b; s_found_above = run; }
            run += h[b];
        }
    }
    {   // clear the table for the next frame
        uint4* z4 = reinterpret_cast<uint4*>(hist + (hi - PER));
#pragma unroll
        for (int v = 0; v < PER / 4; ++v) z4[v] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    if (tid == 0) {
        state[TKS_BIN] = s_found_bin;
        state[TKS_KEFF] = k_eff;
        state[TKS_NEED] = k_eff - s_found_above;
        state[TKS_NSURE] = 0;
        state[TKS_NTIE] = 0;
        state[TKS_VALID] = 0;
        state[TKS_TICKET] = 0;
        *kp_count = (int)k_eff;
    }
}

__device__ __forceinline__ void tk_append(bool mine, unsigned long long key, unsigned* counter, unsigned long long* list,
                                          unsigned lane, unsigned cap) {
    const unsigned m = __ballot_sync(0xffffffffu, mine);
    if (m == 0) return;
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    BSHOT_ASSERT(!mine || base + __popc(m & ((1u << lane) - 1u)) < cap);
    if (mine) list[base + __popc(m & ((1u << lane) - 1u))] = key;
}

__global__ void __launch_bounds__(TK_THREADS)
tk_compact_kernel(const unsigned long long* __restrict__ keys, unsigned n, unsigned* __restrict__ state,
                  unsigned long long* __restrict__ sure, unsigned long long* __restrict__ tie, unsigned sure_cap) {
    const unsigned i = blockIdx.x * TK_THREADS + threadIdx.x, lane = threadIdx.x & 31;
    const unsigned bstar = state[TKS_BIN];
    if (bstar == 0xFFFFFFFFu) return;  // k_eff == 0
    const unsigned long long k = (i < n) ? __ldg(keys + i) : 0ull;
    const unsigned bin = tk_bin(k);
    tk_append(k != 0ull && bin > bstar, k, state + TKS_NSURE, sure, lane, sure_cap);
    tk_append(k != 0ull && bin == bstar, k, state + TKS_NTIE, tie, lane, n);
}

__global__ void __launch_bounds__(TK_THREADS)
tk_rank_kernel(const unsigned* __restrict__ state, const unsigned long long* __restrict__ sure,
               const unsigned long long* __restrict__ tie, const float4* __restrict__ pts, const unsigned* __restrict__ sorted_pos,
               int* __restrict__ kp_flag, int* __restrict__ kp_idx, float* __restrict__ kp_ratio, float4* __restrict__ kp) {
    __shared__ unsigned long long s_keys[TK_TILE];
    __shared__ unsigned s_cnt[16][TK_ITEMS];
    // TK_ITEMS items per CTA; every warp counts one eighth of a tile, its two half warps one sixteenth each (16 items per
    // CTA instead of 32 doubles the number of CTAs with work: K = 10 000 used to keep only half of the GPU busy)
    const unsigned tid = threadIdx.x, slot = tid & (TK_ITEMS - 1), part = tid / TK_ITEMS;
    const unsigned n_sure = state[TKS_NSURE], n_tie = state[TKS_NTIE], need = state[TKS_NEED];
    const unsigned nb_a = (n_sure + TK_ITEMS - 1) / TK_ITEMS, nb_b = (n_tie + TK_ITEMS - 1) / TK_ITEMS;
    for (unsigned ib = blockIdx.x; ib < nb_a + nb_b; ib += gridDim.x) {
        const bool is_a = ib < nb_a;
        const unsigned long long* list = is_a ? sure : tie;
        const unsigned len = is_a ? n_sure : n_tie;
        const unsigned item = (is_a ? ib : ib - nb_a) * TK_ITEMS + slot;
        const unsigned long long x = (item < len) ? list[item] : 0ull;
        unsigned cnt = 0;
        for (unsigned t0 = 0; t0 < len; t0 += TK_TILE) {
            const unsigned tl = min((unsigned)TK_TILE, len - t0);
            __syncthreads();
            for (unsigned t = tid; t < tl; t += TK_THREADS) s_keys[t] = list[t0 + t];
            __syncthreads();
            const unsigned lo = part * (TK_TILE / 16), hi = min(lo + TK_TILE / 16, tl);
            if (is_a) {  // sure keys: position among the sure keys, ascending
                for (unsigned j = lo; j < hi; ++j) cnt += (s_keys[j] < x) ? 1u : 0u;
            } else {     // tie keys: number of larger tie keys
                for (unsigned j = lo; j < hi; ++j) cnt += (s_keys[j] > x) ? 1u : 0u;
            }
        }
        s_cnt[part][slot] = cnt;
        __syncthreads();
        if (part == 0 && item < len) {
            unsigned c = 0;
#pragma unroll
            for (int pp = 0; pp < 16; ++pp) c += s_cnt[pp][slot];
            const bool keep = is_a || c < need;
            if (keep) {
                const unsigned pos = is_a ? need + c : need - 1u - c;
                BSHOT_ASSERT(pos < state[TKS_KEFF]);
                const unsigned idx = ~(unsigned)(x & 0xFFFFFFFFull);
                kp_idx[pos] = (int)idx;
                kp_ratio[pos] = __uint_as_float((unsigned)(x >> 32));
                float4 p = pts[idx];
                p.w = __uint_as_float(idx);
                kp[pos] = p;
                kp_flag[sorted_pos[idx]] = (int)pos;  // keypoint ordinal per cell-sorted position (tiled normals, tilek.cu)
            }
        }
        __syncthreads();
    }
}

static int seg_ratio_warp_launch(Ctx* c, float radius, int max_nn, int sr_type, unsigned ctas, const unsigned* list, const unsigned* list_n) {
    const unsigned n = (unsigned)c->n_points;
#define BSHOT_LAUNCH_SEG(SR)                                                                                                          \
    seg_ratio_kernel<SR><<<ctas, DT_THREADS, 0, c->stream>>>(c->d_grid, c->d_cell_start, c->d_sorted, c->d_pts, n, radius, max_nn, c->d_ratio, \
                                                             c->d_keys, c->d_counters, list, list_n)
    if (sr_type == BSHOT_SR_CV) BSHOT_LAUNCH_SEG(BSHOT_SR_CV);
    else if (sr_type == BSHOT_SR_CVS) BSHOT_LAUNCH_SEG(BSHOT_SR_CVS);
    else BSHOT_LAUNCH_SEG(BSHOT_SR_CVSN);
#undef BSHOT_LAUNCH_SEG
    count_launch(c);
    return check_launch("seg_ratio_kernel");
}

// fuse == 1: also compute the FULL-mode normal of every point from the same neighbourhoods (same radius / max_nn)
// fuse == 2: also keep the nine covariance sums + count of every point's neighbourhood (d_qsums): the REFERENCE-mode normals
//            of the points that become keypoints are then one small eigen-solve each, with no second neighbour search
int detect_seg_ratio(Ctx* c, float radius, int max_nn, int sr_type, int fuse, int gate_top_k) {
    const unsigned n = (unsigned)c->n_points;
    bool fuse_normals = fuse == 1;
    c->fused_normals = false;
    c->fused_sums = false;
    if (n == 0) return BSHOT_OK;
    if (sr_type < 0 || sr_type > 2) { set_error("bad sr_type %d", sr_type); return BSHOT_E_INVALID; }
    if (!(radius > 0.0f)) { set_error("bad radius"); return BSHOT_E_INVALID; }
    mark_unscored_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_pts, c->d_cell_of, n, c->d_ratio, c->d_keys, c->d_rho_hint, radius);
    count_launch(c);
    if (tile_path_ok(c, max_nn)) {
        fuse_normals = fuse_normals && !c->force_warp_path;
        const bool sums = fuse == 2 && !c->force_warp_path && !c->no_deferred_normals;
        // a point the tiles do not answer (fallback list) keeps a NaN count: its normal is searched on its own later
        if (sums) BSHOT_CUDA_TRY(cudaMemsetAsync(c->d_qsums, 0xFF, sizeof(float) * 10 * (size_t)n, c->stream));
        BSHOT_TRY(tile_neighbourhoods(c, sr_type, true, fuse_normals ? 1 : (sums ? 2 : 0), radius, max_nn, nullptr, c->d_normals, sums ? gate_top_k : 0));
        if (sums) { c->fused_sums = true; c->fused_radius = radius; c->fused_max_nn = max_nn; }
        // queries whose tile did not fit (device-side list, usually empty)
        BSHOT_TRY(seg_ratio_warp_launch(c, radius, max_nn, sr_type, (unsigned)c->sm_count * 4u, c->d_fb_list, c->d_nblocks + 1));
        if (fuse_normals) {
            BSHOT_TRY(normals_fallback_list(c, radius, max_nn, nullptr, c->d_normals));
            c->fused_normals = true;
            c->fused_radius = radius;
            c->fused_max_nn = max_nn;
        }
    } else {
        BSHOT_TRY(seg_ratio_warp_launch(c, radius, max_nn, sr_type, (n + DT_WARPS - 1) / DT_WARPS, nullptr, nullptr));
    }
    c->sel_valid = true;
    c->sel_radius = radius;
    c->sel_max_nn = max_nn;
    return BSHOT_OK;
}

#ifdef BSHOT_KNN_STATS
void knn_stats_dump() {
    unsigned long long h[8];
    cudaMemcpyFromSymbol(h, g_knn_stats, sizeof(h));
    fprintf(stderr, "[knn stats] sweeps=%llu rows=%llu cand=%llu insphere=%llu slow=%llu enumerations=%llu enum_rows=%llu\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
    unsigned long long z[8] = {0};
    cudaMemcpyToSymbol(g_knn_stats, z, sizeof(z));
}
#endif

int detect_topk(Ctx* c, int top_k) {
    c->gate_top_k = 0;  // d_kp_ratio is rewritten; frame_extract re-arms the gate for its own parameters
    const unsigned n = (unsigned)c->n_points;
    if (top_k < 0 || (size_t)top_k > c->max_kp) { set_error("top_k %d exceeds max_keypoints %zu", top_k, c->max_kp); return BSHOT_E_CAPACITY; }
    // keypoint ordinal per cell-sorted position, read by the tiled normals (REFERENCE mode): -1 = not a keypoint
    if (n) BSHOT_CUDA_TRY(cudaMemsetAsync(c->d_kp_flag, 0xFF, sizeof(int) * n, c->stream));
    const unsigned hist_ctas = std::max(1u, std::min((n + TK_THREADS * 16 - 1) / (TK_THREADS * 16), (unsigned)c->sm_count));
    tk_hist_kernel<<<hist_ctas, TK_THREADS, 0, c->stream>>>(c->d_keys, n, top_k, c->d_tk_hist, c->d_tk_state, c->d_kp_count);
    if (n) tk_compact_kernel<<<(n + TK_THREADS - 1) / TK_THREADS, TK_THREADS, 0, c->stream>>>(c->d_keys, n, c->d_tk_state, c->d_tk_sure, c->d_tk_tie, (unsigned)c->max_kp);
    tk_rank_kernel<<<(unsigned)c->sm_count * 8u, TK_THREADS, 0, c->stream>>>(c->d_tk_state, c->d_tk_sure, c->d_tk_tie, c->d_pts, c->d_sorted_pos,
                                                                             c->d_kp_flag, c->d_kp_idx, c->d_kp_ratio, c->d_kp);
    count_launch(c, n ? 3 : 2);
#ifdef BSHOT_KNN_STATS
    cudaStreamSynchronize(c->stream);
    knn_stats_dump();
#endif
    c->n_kp = (size_t)top_k;  // upper bound until the host reads d_kp_count
    c->have_kp = true;
    c->kp_from_detector = true;  // d_kp[i].w carries the surface index
    return check_launch("top-K kernels");
}

}  // namespace bshot

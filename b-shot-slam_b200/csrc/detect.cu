// detect.cu -- seg-ratio ("SR") keypoint detector and top-K selection.
//
// Replaces the per-point loop of LidarOdometry::extractKeypoints (src/lidar_odometry.cpp:61-126)
// and the sort / keep-last-K that follows (:131-153).  One warp per point (taken in voxel order so
// neighbouring warps share cache lines): nearest-<=max_nn-inside-R selection (knn.cuh), centroid,
// then the CV / CVS / CVSN score.  Top-K is a single-CTA MSB radix select over 64-bit keys
// (ratio bits << 32 | ~index) followed by an in-shared-memory bitonic sort, so keypoints come out
// in ascending ratio order like the reference's `SegRatio.end()-600 .. end()` slice, with a
// deterministic tie-break (lower point index wins) where std::sort's is unspecified.
#include <stdlib.h>
#include <string.h>

#include "knn.cuh"
#include "stages.h"

namespace bshot {

constexpr int DT_WARPS = 4;
constexpr int DT_THREADS = DT_WARPS * 32;

__global__ void __launch_bounds__(DT_THREADS)
seg_ratio_kernel(const GridParams* __restrict__ gp, const unsigned* __restrict__ cell_start,
                 const float4* __restrict__ sorted, const float4* __restrict__ pts, unsigned n_total, float radius, int max_nn,
                 int sr_type,
                 float* __restrict__ ratio, unsigned long long* __restrict__ keys,
                 unsigned long long* __restrict__ counters, const unsigned* __restrict__ work_list,
                 const unsigned* __restrict__ work_count) {
    __shared__ KnnWarpSmem smem[DT_WARPS];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const GridParams g = *gp;
    KnnWarpSmem& sm = smem[wid];
    // work items: either every binned point (voxel order) or the tile kernel's leftover list
    const unsigned n_items = work_list ? min(*work_count, n_total) : min(__ldg(cell_start + g.ncells), n_total);
    for (unsigned item = blockIdx.x * DT_WARPS + wid; item < n_items; item += gridDim.x * DT_WARPS) {
    const unsigned j = work_list ? work_list[item] : item;
    const float4 q = __ldg(sorted + j);
    const unsigned qi = __float_as_uint(q.w);
    const float nanf_ = __int_as_float(0x7FC00000);
    if (q.x == 0.0f && q.y == 0.0f && q.z == 0.0f) {  // src/lidar_odometry.cpp:63
        if (lane == 0) { ratio[qi] = nanf_; keys[qi] = 0ull; }
        continue;
    }
    RowRange rr;
    // centroid (pcl::computeCentroid, :76): sum in fp64, rounded once
    double sx = 0, sy = 0, sz = 0;
    const KnnResult res = knn_select(g, cell_start, sorted, pts, q, radius, max_nn, sm, lane, rr,
                                     [&](const float4 p) { sx += p.x; sy += p.y; sz += p.z; });
    const float rho = sqrtf(res.rho2) * 1.0001f;
    bool cached = !res.batched;
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
    __syncwarp();
    }  // work items
}

// =====================================================================================================
// Tile path: one LANE per query, one CTA (4 warps) per 32 consecutive points of the voxel-sorted array.
// The 32 points are split into groups that share a row (iy,iz) and an 8-cell x window; a group's
// candidates (the cells within rho of the group) are staged 32 at a time in shared memory -- each warp
// takes every 4th tile -- and lane l of every warp evaluates the staged candidates against query l
// with broadcast LDS.128 reads: no per-candidate search, the candidate loads are amortised over up to
// 32 queries, and the per-query state (128-bin histogram column, counts, centroid sums) lives in
// shared memory and is updated with fire-and-forget shared atomics.  The nearest-<=max_nn selection is
// the same threshold-key construction as knn.cuh.  Groups the tile path cannot take (row rectangle
// larger than the segment list, > 64 Ki candidates, > 32 exact distance duplicates at the threshold)
// go to a leftover list that the warp-per-query kernel above finishes.
constexpr int TL_WARPS = 4;
constexpr int TL_THREADS = TL_WARPS * 32;
constexpr int TL_MAXSEG = 400;
constexpr int TL_MAXB = 256;
constexpr int TL_BINS = 128;
constexpr int TL_LCAP = 32;
constexpr int TL_WINDOW_SHIFT = 3;  // 8-cell x window per group

struct TileSmem {
    SegList<TL_MAXSEG, TL_MAXB> sl;
    float4 tile[TL_WARPS][32];
    union {
        unsigned hist[TL_BINS][32];                // per-query histogram columns (16 KB)
        unsigned long long list[TL_LCAP][32];      // per-query candidate keys of the crossing bin (8 KB)
    } u;
    double sum[3][32];
    unsigned n[32];
    unsigned ln[32];
    int pos[32], neg[32];
};

// every warp sweeps the tiles t = warp, warp + 4, ... of the flattened candidate list
template <typename F>
__device__ __forceinline__ void tile_pass(const float4* __restrict__ sorted, TileSmem& sm, unsigned lane, unsigned wid,
                                          bool active, F&& body) {
    const unsigned total = sm.sl.total;
    float4* tile = sm.tile[wid];
    float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned j0 = wid * 32u;
    if (j0 + lane < total) nxt = __ldg(sorted + seg_lookup(sm.sl, j0 + lane));
    for (; j0 < total; j0 += 32u * TL_WARPS) {
        tile[lane] = nxt;
        __syncwarp();
        const unsigned jn = j0 + 32u * TL_WARPS + lane;
        if (jn < total) nxt = __ldg(sorted + seg_lookup(sm.sl, jn));  // prefetch this warp's next tile
        const unsigned cnt = min(32u, total - j0);
        if (active) {
#pragma unroll 4
            for (unsigned t = 0; t < cnt; ++t) body(tile[t]);
        }
        __syncwarp();
    }
}

__device__ __forceinline__ float tl_bound(float lo, float w, int k) { return fmaf((float)k, w, lo); }

// bin of sqd in [lo, hi) split into TL_BINS bins of width w, consistent with tl_bound()
__device__ __forceinline__ int tl_bin(float sqd, float lo, float w, float inv_w) {
    int b = min(TL_BINS - 1, max(0, (int)((sqd - lo) * inv_w)));
    if (b > 0 && sqd < tl_bound(lo, w, b)) --b;
    else if (b < TL_BINS - 1 && sqd >= tl_bound(lo, w, b + 1)) ++b;
    return b;
}

__global__ void __launch_bounds__(TL_THREADS)
seg_ratio_tile_kernel(const GridParams* __restrict__ gp, const unsigned* __restrict__ cell_start,
                      const float4* __restrict__ sorted, const float4* __restrict__ pts, unsigned n_total, float radius,
                      int max_nn, int sr_type, float* __restrict__ ratio, unsigned long long* __restrict__ keys,
                      unsigned long long* __restrict__ counters, unsigned* __restrict__ leftover,
                      unsigned* __restrict__ leftover_count) {
    __shared__ TileSmem sm;
    const unsigned FULL = 0xffffffffu;
    const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned chunk = blockIdx.x;
    const GridParams g = *gp;
    const unsigned nb = min(__ldg(cell_start + g.ncells), n_total);
    if (chunk * 32u >= nb) return;
    const unsigned j = chunk * 32u + lane;  // every warp holds the same 32 queries
    const bool have = j < nb;
    const float4 q = have ? __ldg(sorted + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    const unsigned qi = __float_as_uint(q.w);
    const float nanf_ = __int_as_float(0x7FC00000);
    bool todo = have;
    if (have && q.x == 0.0f && q.y == 0.0f && q.z == 0.0f) {  // src/lidar_odometry.cpp:63
        if (wid == 0) { ratio[qi] = nanf_; keys[qi] = 0ull; }
        todo = false;
    }
    const int cix = min(max(cell_coord(q.x, g.ox, g.inv_cell), 0), g.nx - 1);
    const int ciy = min(max(cell_coord(q.y, g.oy, g.inv_cell), 0), g.ny - 1);
    const int ciz = min(max(cell_coord(q.z, g.oz, g.inv_cell), 0), g.nz - 1);
    const unsigned gkey = ((unsigned)ciz * g.ny + ciy) * ((unsigned)(g.nx >> TL_WINDOW_SHIFT) + 1u) + (unsigned)(cix >> TL_WINDOW_SHIFT);
    const float R = radius;
    const float R2 = (float)((double)R * (double)R);
    const bool tile_ok = max_nn > 0;
    auto sync = [] { __syncthreads(); };

    auto push_leftover = [&](bool mine) {  // warp 0 only
        const unsigned m = __ballot_sync(FULL, mine);
        if (m == 0 || wid != 0) return;
        unsigned base = 0;
        const int leader = __ffs(m) - 1;
        if ((int)lane == leader) base = atomicAdd(leftover_count, (unsigned)__popc(m));
        base = __shfl_sync(FULL, base, leader);
        if (mine) leftover[base + __popc(m & ((1u << lane) - 1))] = j;
    };

    unsigned remaining = __ballot_sync(FULL, todo);
    while (remaining) {  // uniform across the CTA: every warp computes the same groups
        const int leader = __ffs(remaining) - 1;
        const unsigned lkey = __shfl_sync(FULL, gkey, leader);
        const unsigned gm = __ballot_sync(FULL, todo && gkey == lkey) & remaining;
        remaining &= ~gm;
        bool in_g = (gm >> lane) & 1u;
        if (!tile_ok) {
            push_leftover(in_g);
            continue;
        }
        // ---- grow the shared search region until every query of the group holds >= max_nn -------------
        float rho2 = R2;
        unsigned n = 0;
        bool fail = false;
        for (int m = 1;; ++m) {
            const float g_m = (float)m * g.cell * 0.9999f;
            const bool last = !(g_m < R);
            const float rho = last ? R : g_m;
            rho2 = last ? R2 : __fmul_rn(rho, rho);
            const float pad = rho + 1e-3f * g.cell;
            const int big = 0x3fffffff;
            const int X0 = __reduce_min_sync(FULL, in_g ? max(cell_coord(q.x - pad, g.ox, g.inv_cell), 0) : big);
            const int X1 = __reduce_max_sync(FULL, in_g ? min(cell_coord(q.x + pad, g.ox, g.inv_cell), g.nx - 1) : -1);
            const int Y0 = __reduce_min_sync(FULL, in_g ? max(cell_coord(q.y - pad, g.oy, g.inv_cell), 0) : big);
            const int Y1 = __reduce_max_sync(FULL, in_g ? min(cell_coord(q.y + pad, g.oy, g.inv_cell), g.ny - 1) : -1);
            const int Z0 = __reduce_min_sync(FULL, in_g ? max(cell_coord(q.z - pad, g.oz, g.inv_cell), 0) : big);
            const int Z1 = __reduce_max_sync(FULL, in_g ? min(cell_coord(q.z + pad, g.oz, g.inv_cell), g.nz - 1) : -1);
            const int ys = Y1 - Y0 + 1;
            const int nrows = ys * (Z1 - Z0 + 1);
            if (nrows > TL_MAXSEG || X1 < X0) { fail = true; break; }
            // cheap pre-check from the voxel table: not enough points in the whole region -> grow directly
            __syncthreads();
            if (tid == 0) { sm.sl.nseg = 0; sm.sl.total = 0; }
            if (tid < 32) { sm.n[tid] = 0; }
            __syncthreads();
            for (int r = (int)tid; r < nrows; r += TL_THREADS) {
                const unsigned row = ((unsigned)(Z0 + r / ys) * g.ny + (unsigned)(Y0 + r % ys)) * g.nx;
                const unsigned s = __ldg(cell_start + row + X0), e = __ldg(cell_start + row + X1 + 1);
                if (e > s) {
                    const unsigned slot = atomicAdd(&sm.sl.nseg, 1u);
                    sm.sl.start[slot] = s;
                    sm.sl.off[slot] = e - s;
                    atomicAdd(&sm.sl.total, e - s);
                }
            }
            __syncthreads();
            if (!last && sm.sl.total < (unsigned)max_nn) continue;  // cannot hold max_nn points: next radius
            finish_segments<TL_THREADS>(sm.sl, tid, sync);
            const unsigned total = sm.sl.total;
            for (unsigned i = tid; i < TL_BINS * 32; i += TL_THREADS) (&sm.u.hist[0][0])[i] = 0u;
            __syncthreads();
            const float w = rho2 / (float)TL_BINS, inv_w = (float)TL_BINS / rho2;
            const float lim = rho2;
            unsigned nl = 0;
            tile_pass(sorted, sm, lane, wid, in_g, [&](const float4 p) {
                const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
                if (sqd < lim) {
                    ++nl;
                    atomicAdd(&sm.u.hist[tl_bin(sqd, 0.0f, w, inv_w)][lane], 1u);
                }
            });
            if (in_g && nl) atomicAdd(&sm.n[lane], nl);
            __syncthreads();
            n = sm.n[lane];
            if (last || __all_sync(FULL, !in_g || n >= (unsigned)max_nn)) break;
        }
        if (fail) {
            push_leftover(in_g);
            continue;
        }
        // ---- per-query crossing bin (refined while it holds more than TL_LCAP candidates) ---------------
        const bool sel_all = n <= (unsigned)max_nn;
        bool need_sel = in_g && !sel_all;
        float lo = 0.0f, hi = rho2;
        unsigned below = 0, cntb = 0;
        for (int level = 0; level < 4; ++level) {
            const bool refine = need_sel && (level == 0 || cntb > (unsigned)TL_LCAP);
            if (level > 0) {
                if (!__any_sync(FULL, refine)) break;
                __syncthreads();
                for (unsigned i = tid; i < TL_BINS * 32; i += TL_THREADS) (&sm.u.hist[0][0])[i] = 0u;
                __syncthreads();
                const float w = (hi - lo) / (float)TL_BINS, inv_w = (float)TL_BINS / (hi - lo);
                const float flo = lo, fhi = hi;
                tile_pass(sorted, sm, lane, wid, refine, [&](const float4 p) {
                    const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
                    if (sqd >= flo && sqd < fhi) atomicAdd(&sm.u.hist[tl_bin(sqd, flo, w, inv_w)][lane], 1u);
                });
                __syncthreads();
            }
            if (refine) {  // every warp scans its copy of the column: identical results, no broadcast needed
                const float w = (hi - lo) / (float)TL_BINS;
                unsigned cum = below;
                int b = TL_BINS - 1;
                unsigned hb = 0;
                for (int k = 0; k < TL_BINS; ++k) {
                    const unsigned h = sm.u.hist[k][lane];
                    if (cum + h >= (unsigned)max_nn) { b = k; hb = h; break; }
                    cum += h;
                }
                below = cum;
                cntb = hb;
                const float nlo = tl_bound(lo, w, b);
                const float nhi = (b == TL_BINS - 1) ? hi : tl_bound(lo, w, b + 1);
                lo = nlo;
                hi = nhi;
            }
        }
        // queries whose crossing bin still overflows (exact-distance duplicates) take the exact slow path
        const bool overflow = need_sel && cntb > (unsigned)TL_LCAP;
        push_leftover(overflow);
        if (overflow) { in_g = false; need_sel = false; }
        // ---- collect the crossing bin + centroid of the surely selected --------------------------------
        __syncthreads();
        if (tid < 32) { sm.ln[tid] = 0; sm.pos[tid] = 0; sm.neg[tid] = 0; sm.sum[0][tid] = 0.0; sm.sum[1][tid] = 0.0; sm.sum[2][tid] = 0.0; }
        __syncthreads();
        {
            double sx = 0.0, sy = 0.0, sz = 0.0;
            const float flo = lo, fhi = hi, lim = rho2;
            tile_pass(sorted, sm, lane, wid, in_g, [&](const float4 p) {
                const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
                if (!(sqd < lim)) return;
                if (sel_all || sqd < flo) { sx += p.x; sy += p.y; sz += p.z; }
                else if (sqd < fhi) {
                    const unsigned slot = atomicAdd(&sm.ln[lane], 1u);
                    if (slot < (unsigned)TL_LCAP) sm.u.list[slot][lane] = knn_key(sqd, p.w);
                }
            });
            if (in_g) { atomicAdd(&sm.sum[0][lane], sx); atomicAdd(&sm.sum[1][lane], sy); atomicAdd(&sm.sum[2][lane], sz); }
        }
        __syncthreads();
        double sx = sm.sum[0][lane], sy = sm.sum[1][lane], sz = sm.sum[2][lane];
        unsigned long long thr = (((unsigned long long)__float_as_uint(rho2)) << 32) - 1ull;  // sel_all: sqd < rho2
        int count = (int)n;
        if (need_sel) {  // redundantly in every warp (identical inputs, identical results)
            count = max_nn;
            const unsigned L = min(sm.ln[lane], (unsigned)TL_LCAP);
            const unsigned need = (unsigned)max_nn - below;  // 1..cntb
            for (unsigned e = 0; e < L; ++e) {
                const unsigned long long ke = sm.u.list[e][lane];
                unsigned rank = 0;
                for (unsigned o = 0; o < L; ++o) rank += (sm.u.list[o][lane] < ke) ? 1u : 0u;
                if (rank < need) {
                    const float4 p = __ldg(pts + (unsigned)(ke & 0xFFFFFFFFull));
                    sx += p.x; sy += p.y; sz += p.z;
                    if (rank == need - 1) thr = ke;
                }
            }
        }
        // ---- score -------------------------------------------------------------------------------------
        const float fn = (float)count;
        const float ctx = (float)sx / fn, cty = (float)sy / fn, ctz = (float)sz / fn;
        const float vx = __fsub_rn(q.x, ctx), vy = __fsub_rn(q.y, cty), vz = __fsub_rn(q.z, ctz);  // :79
        float seg;
        if (sr_type == BSHOT_SR_CV) {  // :83-97
            int pos = 0, neg = 0;
            tile_pass(sorted, sm, lane, wid, in_g, [&](const float4 p) {
                const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
                if (knn_key(sqd, p.w) > thr) return;
                const float d = dot3_rn(vx, vy, vz, __fsub_rn(p.x, q.x), __fsub_rn(p.y, q.y), __fsub_rn(p.z, q.z));
                if (d > 0.0f) ++pos;
                else if (d < 0.0f) ++neg;
            });
            if (in_g) { if (pos) atomicAdd(&sm.pos[lane], pos); if (neg) atomicAdd(&sm.neg[lane], neg); }
            __syncthreads();
            const float fp = (float)sm.pos[lane], fq = (float)sm.neg[lane];
            seg = 1.0f - fminf(fp, fq) / fmaxf(fp, fq);
            if (sm.pos[lane] == 0 && sm.neg[lane] == 0) seg = nanf_;
        } else {  // CVS :98-108, CVSN :109-119 (sum reduced across warps through sm.sum[0], zeroed first)
            const float ctn = sqrtf(dot3_rn(vx, vy, vz, vx, vy, vz));
            __syncthreads();
            if (tid < 32) sm.sum[0][tid] = 0.0;
            __syncthreads();
            double sum = 0.0;
            tile_pass(sorted, sm, lane, wid, in_g, [&](const float4 p) {
                const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
                if (knn_key(sqd, p.w) > thr) return;
                const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
                const float dn = sqrtf(dot3_rn(dx, dy, dz, dx, dy, dz));
                if (ctn == 0.0f || dn == 0.0f) return;
                const float d = dot3_rn(vx, vy, vz, dx, dy, dz);
                sum += (sr_type == BSHOT_SR_CVS) ? (double)d : (double)(d / __fmul_rn(ctn, dn));
            });
            if (in_g) atomicAdd(&sm.sum[0][lane], sum);
            __syncthreads();
            seg = fabsf((float)sm.sum[0][lane]) / fn;
        }
        if (wid == 0) {
            if (in_g) {
                ratio[qi] = seg;
                keys[qi] = isnan(seg) ? 0ull : (((unsigned long long)__float_as_uint(seg) << 32) | (unsigned)(~qi));
            }
            const int tot = __reduce_add_sync(FULL, in_g ? count : 0);
            if (lane == 0) atomicAdd(&counters[0], (unsigned long long)tot);
        }
        __syncthreads();
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
    // K-th largest key via MSB-first 8-bit radix select (rank counted from the top). The first pass also
    // counts the valid (non-zero) keys, which fixes k_eff = min(top_k, #valid).
    if (tid == 0) { s_valid = 0; s_fill = 0; s_prefix = 0ull; s_rank = 0; }
    __syncthreads();
    unsigned k_eff = 0;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (unsigned b = tid; b < 256; b += TK_THREADS) hist[b] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        const unsigned long long mask = (shift == 56) ? 0ull : (~0ull << (shift + 8));
        unsigned cv = 0;
        // one CTA reads all keys in every pass: keep 8 independent 8-byte loads in flight per thread
        for (unsigned base = 0; base < n; base += TK_THREADS * 8) {
            unsigned long long kk[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const unsigned i = base + u * TK_THREADS + tid;
                kk[u] = (i < n) ? __ldg(keys + i) : 0ull;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const unsigned i = base + u * TK_THREADS + tid;
                cv += kk[u] != 0ull;
                if (i < n && (kk[u] & mask) == prefix) atomicAdd(&hist[(unsigned)(kk[u] >> shift) & 255u], 1u);
            }
        }
        if (shift == 56) {
            cv = (unsigned)warp_sum((int)cv);
            if ((tid & 31) == 0) atomicAdd(&s_valid, cv);
        }
        __syncthreads();
        if (shift == 56) {
            k_eff = min((unsigned)top_k, s_valid);
            if (tid == 0) { *kp_count = (int)k_eff; s_rank = k_eff ? k_eff - 1 : 0; }
            if (k_eff == 0) return;
            __syncthreads();
        }
        if (tid < 32) {  // warp 0: lane l owns digits 255-8l .. 248-8l (descending), suffix scan across lanes
            unsigned h[8], s = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { h[k] = hist[255 - (tid * 8 + k)]; s += h[k]; }
            unsigned inc = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned up = __shfl_up_sync(0xffffffffu, inc, o);
                if (tid >= (unsigned)o) inc += up;
            }
            const unsigned rank = s_rank;
            unsigned run = inc - s;  // keys in digits above this lane's
            int found = -1;
            unsigned frank = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (found < 0 && rank >= run && rank < run + h[k]) { found = 255 - (int)(tid * 8 + k); frank = rank - run; }
                run += h[k];
            }
            if (found >= 0) {  // exactly one lane
                s_rank = frank;
                s_prefix = prefix | ((unsigned long long)found << shift);
            }
        }
        __syncthreads();
    }
    const unsigned long long kth = s_prefix;  // keys are distinct: exactly k_eff keys are >= kth
    for (unsigned base = 0; base < n; base += TK_THREADS * 8) {
        unsigned long long kk[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const unsigned i = base + u * TK_THREADS + tid;
            kk[u] = (i < n) ? __ldg(keys + i) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (kk[u] >= kth && kk[u] != 0ull) {
                const unsigned slot = atomicAdd(&s_fill, 1u);
                if (slot < sort_cap) sbuf[slot] = kk[u];
            }
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
    // leftover list lives in d_cell_of's neighbour buffer d_qnormals (N x 16 B, free during detection)
    unsigned* leftover = reinterpret_cast<unsigned*>(c->d_qnormals);
    unsigned* leftover_count = reinterpret_cast<unsigned*>(c->d_kp_count) + 2;
    BSHOT_CUDA_TRY(cudaMemsetAsync(leftover_count, 0, sizeof(unsigned), c->stream));
    mark_unbinned_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_cell_of, n, c->d_ratio, c->d_keys);
    static const bool use_tile = [] { const char* e = getenv("BSHOT_DETECTOR"); return e && !strcmp(e, "tile"); }();
    if (use_tile) {
        const unsigned chunks = (n + 31) / 32;
        seg_ratio_tile_kernel<<<chunks, TL_THREADS, 0, c->stream>>>(
            c->d_grid, c->d_cell_start, c->d_sorted, c->d_pts, n, radius, max_nn, sr_type, c->d_ratio, c->d_keys, c->d_counters,
            leftover, leftover_count);
        const unsigned sweep_ctas = std::min((n + DT_WARPS - 1) / DT_WARPS, (unsigned)c->sm_count * 12u);
        seg_ratio_kernel<<<sweep_ctas, DT_THREADS, 0, c->stream>>>(c->d_grid, c->d_cell_start, c->d_sorted, c->d_pts, n, radius, max_nn, sr_type,
                                                                 c->d_ratio, c->d_keys, c->d_counters, leftover, leftover_count);
        count_launch(c, 3);
    } else {
        seg_ratio_kernel<<<(n + DT_WARPS - 1) / DT_WARPS, DT_THREADS, 0, c->stream>>>(
            c->d_grid, c->d_cell_start, c->d_sorted, c->d_pts, n, radius, max_nn, sr_type, c->d_ratio, c->d_keys, c->d_counters, nullptr, nullptr);
        count_launch(c, 2);
    }
    return check_launch("seg_ratio kernels");
}

#ifdef BSHOT_KNN_STATS
void knn_stats_dump() {
    unsigned long long h[8];
    cudaMemcpyFromSymbol(h, g_knn_stats, sizeof(h));
    fprintf(stderr, "[knn stats] attempts=%llu rows=%llu cand=%llu insphere=%llu\n", h[0], h[1], h[2], h[3]);
    unsigned long long z[8] = {0};
    cudaMemcpyToSymbol(g_knn_stats, z, sizeof(z));
}
#endif

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
#ifdef BSHOT_KNN_STATS
    cudaStreamSynchronize(c->stream);
    knn_stats_dump();
#endif
    c->n_kp = (size_t)top_k;  // upper bound until the host reads d_kp_count
    c->have_kp = true;
    return check_launch("topk_kernel");
}

}  // namespace bshot

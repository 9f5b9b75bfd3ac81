// knn.cuh -- "nearest <= max_nn inside radius R" neighbourhood selection, one warp per query.
//
// This is pcl::KdTreeFLANN::radiusSearch(p, R, idx, sqd, max_nn) as the reference calls it
// (src/lidar_odometry.cpp:70, include/bshot_bits.h:68): the hits inside the
// radius, and when there are more than max_nn of them the max_nn NEAREST, ordered by
// (fp32 squared distance, point index).  The warp never materialises the list; it computes a
// THRESHOLD KEY (sqd bits << 32 | index) such that the selected set is {key <= threshold}:
//   see knn_select() below.
// Callers then sweep the SAME shared-memory segment list with `knn_selected()` as the predicate.
#pragma once
#include "nbr.cuh"

namespace bshot {

constexpr int KN_MAXSEG = 400;   // rows of the largest query rectangle kept per warp
constexpr int KN_BINS = 256;
constexpr int KN_LIST = 256;
constexpr int KN_MAXB = 256;     // batch table covers 8192 candidates per query

#ifdef BSHOT_KNN_STATS
__device__ unsigned long long g_knn_stats[8];
#endif

struct KnnWarpSmem {
    SegList<KN_MAXSEG, KN_MAXB> sl;
    unsigned hist[KN_BINS];
    unsigned long long list[KN_LIST];
    unsigned list_n;
    unsigned long long thr;
};

struct KnnResult {
    float rho2;               // squared search radius the segment list was built for
    unsigned long long thr;   // selected <=> sqd < rho2 && key <= thr
    int count;                // size of the selected set
    bool batched;             // segment list does not cover all rows (callers must re-batch)
};

__device__ __forceinline__ unsigned long long knn_key(float sqd, float w) {
    return ((unsigned long long)__float_as_uint(sqd) << 32) | __float_as_uint(w);
}

// iterate all candidates of the query sphere (p, rho); f(float4 point). Rebuilds the list per batch
// only when the row rectangle does not fit (never for the default cell / radius ratio).
template <typename F>
__device__ __forceinline__ void knn_for_each(const GridParams& g, const unsigned* __restrict__ cell_start,
                                             const float4* __restrict__ sorted, const float4& q, float rho,
                                             const RowRange& rr, KnnWarpSmem& sm, unsigned lane, bool& cached, F&& f) {
    auto sync = [] { __syncwarp(); };
    for (int row0 = 0; row0 < rr.nrows; row0 += KN_MAXSEG) {
        if (!(cached && rr.nrows <= KN_MAXSEG)) {
            build_segments<32, KN_MAXSEG, KN_MAXB>(g, cell_start, q.x, q.y, q.z, rho, rr, row0, sm.sl, lane, sync);
            cached = true;
        }
        const unsigned total = sm.sl.total;
        // 4 independent candidate loads in flight per lane before the first use
        for (unsigned j = lane; j < total; j += 32 * 4) {
            float4 p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned ju = j + 32u * u;
                p[u] = __ldg(sorted + seg_lookup(sm.sl, ju < total ? ju : j));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (j + 32u * u < total) f(p[u]);
        }
        __syncwarp();
    }
}

// All 32 lanes call.  Selects the nearest <= max_nn points inside radius R of q and calls acc(p) exactly
// once (on some lane) for every selected point; the caller reduces its accumulators across the warp.
// On return sm.sl holds the segment list for radius sqrt(res.rho2) (valid for re-use iff !res.batched),
// `rr_out` the matching row rectangle and res.thr the threshold key for further sweeps.
//   1. probe rho = 2 cells; predict the radius that holds max_nn points from the local density and grow
//      from there (one 256-bin sqd histogram sweep per attempt; a cube of m cells around the query's cell
//      contains every point closer than m * cell)
//   2. crossing bin of the histogram (re-histogrammed inside the bin while it holds > KN_LIST candidates)
//   3. one sweep: accumulate everything below the crossing bin, collect the bin; rank the short list by
//      key, accumulate its first (max_nn - below) entries (re-read from the original-order array `pts`)
template <typename Acc>
__device__ __forceinline__ KnnResult knn_select(const GridParams& g, const unsigned* __restrict__ cell_start,
                                                const float4* __restrict__ sorted, const float4* __restrict__ pts,
                                                const float4& q, float R, int max_nn, KnnWarpSmem& sm, unsigned lane,
                                                RowRange& rr_out, Acc&& acc) {
    KnnResult res;
    const float R2 = (float)((double)R * (double)R);
    float rho = R, rho2 = R2;
    int n = 0;
    bool cached = false;
    RowRange rr;
    auto sync = [] { __syncwarp(); };
    const int M = max(1, (int)ceilf(R * g.inv_cell)) + 1;  // rho(M) >= R: the loop always terminates at m == M
    // ---- 1. grow the sphere ---------------------------------------------------------------------
    for (int m = min(2, M);;) {
        const float g_m = (float)m * g.cell * 0.9999f;
        const bool last = (max_nn <= 0) || m >= M || !(g_m < R);
        rho = last ? R : g_m;
        rho2 = last ? R2 : __fmul_rn(rho, rho);
        rr = row_range(g, q.y, q.z, rho);
        cached = false;
        unsigned total = 0xFFFFFFFFu;
        if (rr.nrows <= KN_MAXSEG) {
            // cheap necessary condition first: the candidate rows must hold at least max_nn points
            total = enumerate_segments<32>(g, cell_start, q.x, q.y, q.z, rho, rr, 0, sm.sl, lane, sync);
            if (!last && total < (unsigned)max_nn) {
                // surface-like density: count ~ r^2  ->  radius that should hold 1.4 * max_nn candidates
                const float f = sqrtf(1.4f * (float)max_nn / (float)max(total, 1u));
                m = min(M, max(m + 1, (int)ceilf((float)m * f)));
                continue;
            }
            finish_segments<32>(sm.sl, lane, sync);
            cached = true;
        }
        for (unsigned b = lane; b < KN_BINS; b += 32) sm.hist[b] = 0;
        __syncwarp();
        const float scale = (float)KN_BINS / rho2;
        int cnt = 0;
        knn_for_each(g, cell_start, sorted, q, rho, rr, sm, lane, cached, [&](const float4 p) {
            const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
            if (sqd < rho2) {
                ++cnt;
                atomicAdd(&sm.hist[min(KN_BINS - 1, (int)(sqd * scale))], 1u);
            }
        });
        n = warp_sum(cnt);
#ifdef BSHOT_KNN_STATS
        if (lane == 0) { atomicAdd(&g_knn_stats[0], 1ull); atomicAdd(&g_knn_stats[1], (unsigned long long)rr.nrows); atomicAdd(&g_knn_stats[2], (unsigned long long)sm.sl.total); atomicAdd(&g_knn_stats[3], (unsigned long long)n); }
#endif
        if (last || n >= max_nn) break;
        const float f = sqrtf(1.15f * (float)max_nn / (float)max(n, 1));
        m = min(M, max(m + 1, (int)ceilf((float)m * f)));
    }
    rr_out = rr;
    res.rho2 = rho2;
    res.batched = rr.nrows > KN_MAXSEG;
    if (max_nn <= 0 || n <= max_nn) {  // everything inside the sphere is selected
        res.thr = (((unsigned long long)__float_as_uint(rho2)) << 32) - 1ull;
        res.count = n;
        knn_for_each(g, cell_start, sorted, q, rho, rr, sm, lane, cached, [&](const float4 p) {
            if (sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z) < rho2) acc(p);
        });
        return res;
    }
    // ---- 2. crossing bin (the level-0 histogram is already in sm.hist) ----------------------------------
    float lo = 0.0f, hi = rho2;
    int below = 0;  // selected-for-sure elements with sqd < lo
    for (int iter = 0; iter < 8; ++iter) {
        const float scale = (float)KN_BINS / (hi - lo);
        if (iter > 0) {
            for (unsigned b = lane; b < KN_BINS; b += 32) sm.hist[b] = 0;
            __syncwarp();
            knn_for_each(g, cell_start, sorted, q, rho, rr, sm, lane, cached, [&](const float4 p) {
                const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
                if (sqd >= lo && sqd < hi) atomicAdd(&sm.hist[min(KN_BINS - 1, (int)((sqd - lo) * scale))], 1u);
            });
        }
        // locate the crossing bin: lane l owns bins 8l..8l+7
        unsigned h[8], s = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { h[k] = sm.hist[lane * 8 + k]; s += h[k]; }
        unsigned inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += up;
        }
        unsigned run = (unsigned)below + inc - s;
        int found_bin = -1;
        unsigned found_below = 0, found_cnt = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (found_bin < 0 && run < (unsigned)max_nn && run + h[k] >= (unsigned)max_nn) {
                found_bin = (int)lane * 8 + k;
                found_below = run;
                found_cnt = h[k];
            }
            run += h[k];
        }
        const unsigned who = __ballot_sync(0xffffffffu, found_bin >= 0);
        const int src = __ffs(who) - 1;  // exactly one lane finds it (n > max_nn)
        const int bin = __shfl_sync(0xffffffffu, found_bin, src);
        const unsigned bbelow = __shfl_sync(0xffffffffu, found_below, src);
        const unsigned bcnt = __shfl_sync(0xffffffffu, found_cnt, src);
        const float blo = lo, bhi = hi, bscale = scale;
        auto bin_of = [&](float sqd) { return min(KN_BINS - 1, (int)((sqd - blo) * bscale)); };
        if (bcnt <= KN_LIST || iter == 7) {
            // ---- 3. accumulate below the bin, collect the bin, rank, accumulate the rest ---------------
            if (lane == 0) sm.list_n = 0;
            __syncwarp();
            knn_for_each(g, cell_start, sorted, q, rho, rr, sm, lane, cached, [&](const float4 p) {
                const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
                if (!(sqd < bhi)) return;
                const int b = (sqd >= blo) ? bin_of(sqd) : -1;
                if (b < bin) acc(p);
                else if (b == bin) {
                    const unsigned slot = atomicAdd(&sm.list_n, 1u);
                    if (slot < KN_LIST) sm.list[slot] = knn_key(sqd, p.w);
                }
            });
            const unsigned ln = min(sm.list_n, (unsigned)KN_LIST);
            const unsigned need = (unsigned)max_nn - bbelow;  // 1..bcnt
            // fallback threshold (only reachable for > KN_LIST exact distance duplicates)
            if (lane == 0) sm.thr = (unsigned long long)__float_as_uint(hi) << 32;
            __syncwarp();
            for (unsigned e = lane; e < ln; e += 32) {
                const unsigned long long ke = sm.list[e];
                unsigned rank = 0;
                for (unsigned o = 0; o < ln; ++o) rank += (sm.list[o] < ke) ? 1u : 0u;
                if (rank < need) {
                    acc(__ldg(pts + (unsigned)(ke & 0xFFFFFFFFull)));
                    if (rank == need - 1) sm.thr = ke;
                }
            }
            __syncwarp();
            res.thr = sm.thr;
            res.count = max_nn;
            return res;
        }
        // narrow to the crossing bin and histogram again; `below` is recounted exactly for the new bound
        const float w = (hi - lo) / (float)KN_BINS;
        const float nlo = lo + w * (float)bin, nhi = lo + w * (float)(bin + 1);
        lo = fmaxf(lo, nlo - w * 1e-3f);
        hi = fminf(hi, nhi + w * 1e-3f);
        int cb = 0;
        const float flo = lo;
        knn_for_each(g, cell_start, sorted, q, rho, rr, sm, lane, cached, [&](const float4 p) {
            if (sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z) < flo) ++cb;
        });
        below = warp_sum(cb);
    }
    res.thr = (((unsigned long long)__float_as_uint(rho2)) << 32) - 1ull;  // unreachable
    res.count = n;
    return res;
}

__device__ __forceinline__ bool knn_selected(const KnnResult& r, float sqd, float w) {
    return knn_key(sqd, w) <= r.thr;  // thr < (rho2 bits << 32) by construction
}

}  // namespace bshot

// knn.cuh -- "nearest <= max_nn inside radius R" neighbourhood selection, one warp per query.
//
// This is pcl::KdTreeFLANN::radiusSearch(p, R, idx, sqd, max_nn) as the reference calls it
// (src/lidar_odometry.cpp:70, include/bshot_bits.h:68): the hits inside the
// radius, and when there are more than max_nn of them the max_nn NEAREST, ordered by
// (fp32 squared distance, point index).  The warp never materialises the list; it computes a
// THRESHOLD KEY (sqd bits << 32 | index) such that the selected set is {key <= threshold}:
//   1. grow a search sphere rho = cell, 2 cell, 4 cell ... (<= R) until it holds >= max_nn points
//      (a cube of m cells around the query's cell contains every point closer than m * cell)
//   2. 256-bin histogram of sqd in shared memory -> the bin where the cumulative count crosses max_nn
//      (re-histogrammed inside that bin while it holds more than KN_LIST candidates)
//   3. collect that bin's candidates, rank them by key, pick the (max_nn - below)-th
// Callers then sweep the SAME shared-memory segment list with `selected()` as the predicate.
#pragma once
#include "nbr.cuh"

namespace bshot {

constexpr int KN_MAXSEG = 400;   // rows of the largest query rectangle kept per warp
constexpr int KN_BINS = 256;
constexpr int KN_LIST = 256;
constexpr int KN_MAXB = 256;     // batch table covers 8192 candidates per query

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

// All 32 lanes call. On return sm.sl holds the segment list for radius sqrt(res.rho2) (valid for
// re-use iff !res.batched) and `rr_out` the matching row rectangle.
__device__ __forceinline__ KnnResult knn_select(const GridParams& g, const unsigned* __restrict__ cell_start,
                                                const float4* __restrict__ sorted, const float4& q, float R, int max_nn,
                                                KnnWarpSmem& sm, unsigned lane, RowRange& rr_out) {
    KnnResult res;
    const float R2 = (float)((double)R * (double)R);
    float rho = R, rho2 = R2;
    int n = 0;
    bool cached = false;
    RowRange rr;
    // ---- 1. grow the sphere: rho = cell, 2 cell, 3 cell ... until it holds >= max_nn points ----------
    auto sync = [] { __syncwarp(); };
    for (int m = 1;; ++m) {
        const float g_m = (float)m * g.cell * 0.9999f;
        const bool last = (max_nn <= 0) || !(g_m < R);
        rho = last ? R : g_m;
        rho2 = last ? R2 : __fmul_rn(rho, rho);
        rr = row_range(g, q.y, q.z, rho);
        cached = false;
        if (rr.nrows <= KN_MAXSEG) {
            // cheap necessary condition first: the candidate rows must hold at least max_nn points
            const unsigned total = enumerate_segments<32>(g, cell_start, q.x, q.y, q.z, rho, rr, 0, sm.sl, lane, sync);
            if (!last && total < (unsigned)max_nn) continue;
            finish_segments<32>(sm.sl, lane, sync);
            cached = true;
        }
        int cnt = 0;
        knn_for_each(g, cell_start, sorted, q, rho, rr, sm, lane, cached, [&](const float4 p) {
            if (sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z) < rho2) ++cnt;
        });
        n = warp_sum(cnt);
        if (last || n >= max_nn) break;
    }
    rr_out = rr;
    res.rho2 = rho2;
    res.batched = rr.nrows > KN_MAXSEG;
    if (max_nn <= 0 || n <= max_nn) {
        res.thr = 0xFFFFFFFFFFFFFFFFull;
        res.count = n;
        return res;
    }
    // ---- 2. histogram refinement ----------------------------------------------------------------
    float lo = 0.0f, hi = rho2;
    int below = 0;  // selected-for-sure elements with sqd < lo
    for (int iter = 0; iter < 8; ++iter) {
        for (unsigned b = lane; b < KN_BINS; b += 32) sm.hist[b] = 0;
        __syncwarp();
        const float scale = (float)KN_BINS / (hi - lo);
        knn_for_each(g, cell_start, sorted, q, rho, rr, sm, lane, cached, [&](const float4 p) {
            const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
            if (sqd >= lo && sqd < hi) {
                const int b = min(KN_BINS - 1, (int)((sqd - lo) * scale));
                atomicAdd(&sm.hist[b], 1u);
            }
        });
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
        // the bin's value range, using the same arithmetic as the binning above
        const float blo = lo, bscale = scale;
        below = (int)bbelow;
        auto in_bin = [&](float sqd) {
            return sqd >= blo && sqd < hi && min(KN_BINS - 1, (int)((sqd - blo) * bscale)) == bin;
        };
        if (bcnt <= KN_LIST || iter == 7) {
            // ---- 3. collect + rank ------------------------------------------------------------------
            if (lane == 0) sm.list_n = 0;
            __syncwarp();
            knn_for_each(g, cell_start, sorted, q, rho, rr, sm, lane, cached, [&](const float4 p) {
                const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
                if (in_bin(sqd)) {
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
                if (rank == need - 1) sm.thr = ke;
            }
            __syncwarp();
            res.thr = sm.thr;
            res.count = max_nn;
            return res;
        }
        // narrow to the crossing bin and histogram again
        const float w = (hi - lo) / (float)KN_BINS;
        const float nlo = lo + w * (float)bin, nhi = lo + w * (float)(bin + 1);
        // keep the bin membership consistent with in_bin(): widen by one ulp-ish margin
        lo = fmaxf(lo, nlo - w * 1e-3f);
        hi = fminf(hi, nhi + w * 1e-3f);
        // elements of other bins that fall into the widened margin are counted again below, so
        // recompute `below` exactly for the new lower bound
        int cb = 0;
        const float flo = lo;
        knn_for_each(g, cell_start, sorted, q, rho, rr, sm, lane, cached, [&](const float4 p) {
            if (sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z) < flo) ++cb;
        });
        below = warp_sum(cb);
    }
    res.thr = 0xFFFFFFFFFFFFFFFFull;  // unreachable
    res.count = n;
    return res;
}

__device__ __forceinline__ bool knn_selected(const KnnResult& r, float sqd, float w) {
    return sqd < r.rho2 && knn_key(sqd, w) <= r.thr;
}

}  // namespace bshot

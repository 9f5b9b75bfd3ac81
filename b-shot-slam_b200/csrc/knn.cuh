// knn.cuh -- "nearest <= max_nn inside radius R" neighbourhood selection, one warp per query.
//
// This is pcl::KdTreeFLANN::radiusSearch(p, R, idx, sqd, max_nn) as the reference calls it
// (src/lidar_odometry.cpp:70, include/bshot_bits.h:68): the hits inside the
// radius, and when there are more than max_nn of them the max_nn NEAREST, ordered by
// (fp32 squared distance, point index).  The warp never materialises the sorted list; it computes a
// THRESHOLD KEY (sqd bits << 32 | index) such that the selected set is {key <= threshold}:
//   see knn_select() below.
// Candidates: the rows of the voxel table that the search sphere touches are expanded ONCE per sphere
// into an explicit list of indices into the cell-sorted array (shared memory, KN_CAP entries per warp),
// so every later sweep is `sorted[idx[j]]` -- one LDS + one coalesced-ish 16-byte load per candidate,
// no per-candidate segment search.  Spheres with more than KN_CAP candidates (very dense spots) fall
// back to the segment list + batch table of nbr.cuh, which aliases the same shared memory.
#pragma once
#include "nbr.cuh"

namespace bshot {

constexpr int KN_MAXSEG = 400;   // slow path: rows of the largest query rectangle kept per warp
constexpr int KN_MAXB = 256;     // slow path: batch table covers 8192 candidates per query
constexpr int KN_CAP = 1024;     // fast path: explicit candidate list
constexpr int KN_BINS = 256;
constexpr int KN_LIST = 224;

#ifdef BSHOT_KNN_STATS
__device__ unsigned long long g_knn_stats[8];
#endif

struct KnnWarpSmem {
    union {
        unsigned idx[KN_CAP];               // fast path: positions in the cell-sorted array
        SegList<KN_MAXSEG, KN_MAXB> sl;     // slow path
    } u;
    unsigned hist[KN_BINS];
    unsigned long long list[KN_LIST];
    unsigned list_n;
    unsigned long long thr;
};
static_assert(sizeof(SegList<KN_MAXSEG, KN_MAXB>) <= sizeof(unsigned) * KN_CAP, "slow path must fit under the index list");

// where the candidates of the current sphere live (warp-uniform)
struct KnnIter {
    RowRange rr;
    float rho;        // radius the candidates were enumerated for
    unsigned total;   // number of candidates (list mode)
    bool list;        // true: sm.u.idx[0..total) ; false: segment list (sm.u.sl), rebuilt when !cached
    bool cached;
};

struct KnnResult {
    float rho2;               // squared search radius the candidates were enumerated for
    unsigned long long thr;   // selected <=> sqd < rho2 && key <= thr
    int count;                // size of the selected set
    int n_in;                 // points inside the final sphere (>= count)
    int m;                    // sphere size (cells) that succeeded: warm start for a nearby query
    KnnIter it;               // candidate set for further sweeps by the caller
};

__device__ __forceinline__ unsigned long long knn_key(float sqd, float w) {
    return ((unsigned long long)__float_as_uint(sqd) << 32) | __float_as_uint(w);
}

// Enumerate the row segments of the sphere (q, rho) and expand them into sm.u.idx (entries beyond
// KN_CAP are dropped but still counted).  Returns the number of candidates.  All 32 lanes call.
__device__ __forceinline__ unsigned knn_expand(const GridParams& g, const unsigned* __restrict__ cell_start,
                                               const float4& q, float rho, const RowRange& rr, KnnWarpSmem& sm,
                                               unsigned lane) {
    unsigned base = 0;
    __syncwarp();
    for (int r0 = 0; r0 < rr.nrows; r0 += 32) {
        const int r = r0 + (int)lane;
        unsigned s = 0, len = 0;
        if (r < rr.nrows) {
            int iy, iz;
            row_coords(rr, r, iy, iz);
            unsigned e;
            if (row_segment(g, cell_start, q.x, q.y, q.z, rho, iy, iz, s, e)) len = e - s;
        }
        unsigned inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += up;
        }
        const unsigned off = base + inc - len;
        const unsigned room = (off < (unsigned)KN_CAP) ? (unsigned)KN_CAP - off : 0u;
        const unsigned wr = min(len, room);
        // lane-divergent fill, 4 entries per trip (segments are short: a few points per row on lidar data)
        for (unsigned i = 0; i < wr; i += 4) {
            unsigned* d = sm.u.idx + off + i;
            const unsigned v = s + i;
            d[0] = v;
            if (i + 1 < wr) d[1] = v + 1;
            if (i + 2 < wr) d[2] = v + 2;
            if (i + 3 < wr) d[3] = v + 3;
        }
        base += __shfl_sync(0xffffffffu, inc, 31);
    }
    __syncwarp();
    return base;
}

// iterate all candidates of the current sphere; f(float4 point).
template <typename F>
__device__ __forceinline__ void knn_for_each(const GridParams& g, const unsigned* __restrict__ cell_start,
                                             const float4* __restrict__ sorted, const float4& q, KnnIter& it,
                                             KnnWarpSmem& sm, unsigned lane, F&& f) {
    if (it.list) {
        const unsigned total = it.total;
        // 4 independent candidate loads in flight per lane before the first use
        for (unsigned j = lane; j < total; j += 32 * 4) {
            float4 p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned ju = j + 32u * u;
                p[u] = __ldg(sorted + sm.u.idx[ju < total ? ju : j]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (j + 32u * u < total) f(p[u]);
        }
        __syncwarp();
        return;
    }
    auto sync = [] { __syncwarp(); };
    for (int row0 = 0; row0 < it.rr.nrows; row0 += KN_MAXSEG) {
        if (!(it.cached && it.rr.nrows <= KN_MAXSEG)) {
            build_segments<32, KN_MAXSEG, KN_MAXB>(g, cell_start, q.x, q.y, q.z, it.rho, it.rr, row0, sm.u.sl, lane, sync);
            it.cached = true;
        }
        const unsigned total = sm.u.sl.total;
#pragma unroll 1
        for (unsigned j = lane; j < total; j += 32 * 2) {
            const unsigned j1 = j + 32u;
            const float4 p0 = __ldg(sorted + seg_lookup(sm.u.sl, j));
            const float4 p1 = __ldg(sorted + seg_lookup(sm.u.sl, j1 < total ? j1 : j));
            f(p0);
            if (j1 < total) f(p1);
        }
        __syncwarp();
    }
}

// All 32 lanes call.  Selects the nearest <= max_nn points inside radius R of q and calls acc(p) exactly
// once (on some lane) for every selected point; the caller reduces its accumulators across the warp.
// On return res.it describes the candidate set (still in shared memory) for further sweeps with
// knn_selected() as the predicate.  `m_hint` > 0 starts the sphere growth at that many cells (the size
// that worked for a nearby query) instead of the 2-cell probe.
//   1. grow the sphere: candidates are enumerated first (cheap) and a sweep is only spent when the rows
//      can hold max_nn points; the radius is predicted from the local density (count ~ r^2 on surfaces);
//      one sweep counts the points inside the sphere and fills a 256-bin sqd histogram
//   2. crossing bin of the histogram (re-histogrammed inside the bin while it holds > KN_LIST candidates)
//   3. one sweep: accumulate everything below the crossing bin, collect the bin; rank the short list by
//      key, accumulate its first (max_nn - below) entries (re-read from the original-order array `pts`)
template <typename Acc>
__device__ __forceinline__ KnnResult knn_select(const GridParams& g, const unsigned* __restrict__ cell_start,
                                                const float4* __restrict__ sorted, const float4* __restrict__ pts,
                                                const float4& q, float R, int max_nn, int m_hint, KnnWarpSmem& sm,
                                                unsigned lane, Acc&& acc) {
    KnnResult res;
    KnnIter& it = res.it;
    const float R2 = (float)((double)R * (double)R);
    float rho2 = R2;
    int n = 0;
    const int M = max(1, (int)ceilf(R * g.inv_cell)) + 1;  // rho(M) >= R: the loop always terminates at m == M
    int m = (m_hint > 0) ? min(m_hint, M) : min(2, M);
    // ---- 1. grow the sphere ---------------------------------------------------------------------
    for (;;) {
        const float g_m = (float)m * g.cell * 0.9999f;
        const bool last = (max_nn <= 0) || m >= M || !(g_m < R);
        it.rho = last ? R : g_m;
        rho2 = last ? R2 : __fmul_rn(it.rho, it.rho);
        it.rr = row_range(g, q.y, q.z, it.rho);
        it.total = knn_expand(g, cell_start, q, it.rho, it.rr, sm, lane);
        if (!last && it.total < (unsigned)max_nn) {
            // cheap necessary condition failed (the candidate rows hold fewer than max_nn points).
            // surface-like density: count ~ r^2  ->  radius that should hold 1.4 * max_nn candidates
            const float f = sqrtf(1.4f * (float)max_nn / (float)max(it.total, 1u));
            m = min(M, max(m + 1, (int)ceilf((float)m * f)));
            continue;
        }
        it.list = it.total <= (unsigned)KN_CAP;
        it.cached = false;
        for (unsigned b = lane; b < KN_BINS; b += 32) sm.hist[b] = 0;
        __syncwarp();
        const float scale = (float)KN_BINS / rho2;
        int cnt = 0;
        knn_for_each(g, cell_start, sorted, q, it, sm, lane, [&](const float4 p) {
            const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
            if (sqd < rho2) {
                ++cnt;
                atomicAdd(&sm.hist[min(KN_BINS - 1, (int)(sqd * scale))], 1u);
            }
        });
        n = warp_sum(cnt);
#ifdef BSHOT_KNN_STATS
        if (lane == 0) { atomicAdd(&g_knn_stats[0], 1ull); atomicAdd(&g_knn_stats[1], (unsigned long long)it.rr.nrows); atomicAdd(&g_knn_stats[2], (unsigned long long)it.total); atomicAdd(&g_knn_stats[3], (unsigned long long)n); atomicAdd(&g_knn_stats[4], it.list ? 0ull : 1ull); }
#endif
        if (last || n >= max_nn) break;
        const float f = sqrtf(1.15f * (float)max_nn / (float)max(n, 1));
        m = min(M, max(m + 1, (int)ceilf((float)m * f)));
    }
    res.rho2 = rho2;
    res.m = m;
    res.n_in = n;
    if (max_nn <= 0 || n <= max_nn) {  // everything inside the sphere is selected
        res.thr = (((unsigned long long)__float_as_uint(rho2)) << 32) - 1ull;
        res.count = n;
        knn_for_each(g, cell_start, sorted, q, it, sm, lane, [&](const float4 p) {
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
            int cb = 0;
            const float flo = lo, fhi = hi;
            knn_for_each(g, cell_start, sorted, q, it, sm, lane, [&](const float4 p) {
                const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
                if (sqd < flo) ++cb;  // `below` is recounted exactly for the narrowed bound
                else if (sqd < fhi) atomicAdd(&sm.hist[min(KN_BINS - 1, (int)((sqd - flo) * scale))], 1u);
            });
            below = warp_sum(cb);
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
            knn_for_each(g, cell_start, sorted, q, it, sm, lane, [&](const float4 p) {
                const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
                if (!(sqd < bhi)) return;
                const int b = (sqd >= blo) ? bin_of(sqd) : -1;
                if (b < bin) acc(p);
                else if (b == bin) {
                    const unsigned slot = atomicAdd(&sm.list_n, 1u);
                    if (slot < KN_LIST) sm.list[slot] = knn_key(sqd, p.w);
                }
            });
            __syncwarp();
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
        // narrow to the crossing bin and histogram again
        const float w = (hi - lo) / (float)KN_BINS;
        const float nlo = lo + w * (float)bin, nhi = lo + w * (float)(bin + 1);
        lo = fmaxf(lo, nlo - w * 1e-3f);
        hi = fminf(hi, nhi + w * 1e-3f);
    }
    res.thr = (((unsigned long long)__float_as_uint(rho2)) << 32) - 1ull;  // unreachable
    res.count = n;
    return res;
}

__device__ __forceinline__ bool knn_selected(const KnnResult& r, float sqd, float w) {
    return knn_key(sqd, w) <= r.thr;  // thr < (rho2 bits << 32) by construction
}

}  // namespace bshot

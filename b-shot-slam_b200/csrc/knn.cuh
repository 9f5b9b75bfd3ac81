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
#ifndef BSHOT_KN_DEPTH
#define BSHOT_KN_DEPTH 4
#endif
#ifndef BSHOT_KN_WLO
#define BSHOT_KN_WLO 1.4f
#endif
#ifndef BSHOT_KN_WHI
#define BSHOT_KN_WHI 2.4f
#endif
#ifndef BSHOT_KN_PAD
#define BSHOT_KN_PAD 0.5f  // candidate-count model: count ~ (rho + PAD * row thickness)^2 (rows are taken whole in y and z)
#endif
constexpr int KN_DEPTH = BSHOT_KN_DEPTH;
constexpr int KN_BINS = 256;
constexpr int KN_LIST = 224;
constexpr int KN_SEGCAP = 2 * KN_LIST;  // fast path: non-empty row segments per sphere (staging aliases `list` and `hist`)

#ifdef BSHOT_KNN_STATS
__device__ unsigned long long g_knn_stats[8];
#endif

struct KnnWarpSmem {
    union {
        unsigned idx[KN_CAP];               // fast path: positions in the cell-sorted array
        SegList<KN_MAXSEG, KN_MAXB> sl;     // slow path
    } u;
    union {
        unsigned hist[KN_BINS];             // sqd histogram
        struct { unsigned heads[32]; unsigned short off[KN_SEGCAP]; } x;  // expansion: head masks of the list chunks, segment offsets
    } h;
    union {
        unsigned long long list[KN_LIST];   // crossing-bin keys
        unsigned seg_start[KN_SEGCAP];      // expansion: first point of every non-empty row segment
    } v;
    unsigned list_n;
    unsigned long long thr;
};
static_assert(sizeof(SegList<KN_MAXSEG, KN_MAXB>) <= sizeof(unsigned) * KN_CAP, "slow path must fit under the index list");
static_assert(KN_CAP <= 32 * 32, "one head-mask word per lane");
static_assert(32 * 4 + KN_SEGCAP * 2 <= KN_BINS * 4 && KN_CAP < 65536, "expansion staging must fit under the histogram");

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
    KnnIter it;               // candidate set for further sweeps by the caller
};

__device__ __forceinline__ unsigned long long knn_key(float sqd, float w) {
    return ((unsigned long long)__float_as_uint(sqd) << 32) | __float_as_uint(w);
}

// Enumerate the row segments of the sphere (q, rho) and, when there are at least `fill_lo` and at most
// `fill_hi` (<= KN_CAP) candidates in at most KN_SEGCAP non-empty segments, expand them into sm.u.idx.  Returns the
// number of candidates; `filled` says whether sm.u.idx holds them.  All 32 lanes call.
//   1. rows -> (start, length) of the non-empty segments, compacted with a ballot rank (row order)
//   2. exclusive scan of the lengths
//   3. head mask: bit j of word c set iff a segment starts at candidate 32 c + j
//   4. balanced fill: candidate j belongs to segment #(heads at or before j); every STS writes 32 entries
__device__ __forceinline__ unsigned knn_expand(const GridParams& g, const unsigned* __restrict__ cell_start,
                                               const float4& q, float rho, const RowRange& rr, unsigned fill_lo,
                                               unsigned fill_hi, KnnWarpSmem& sm, unsigned lane, bool& filled) {
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned nseg = 0, part = 0;
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
        const unsigned m = __ballot_sync(0xffffffffu, len > 0);
        const unsigned k = nseg + __popc(m & lt_mask);
        // lengths above 65535 cannot be listed anyway (total > KN_CAP): clamp for the u16 staging
        if (len > 0 && k < (unsigned)KN_SEGCAP) { sm.v.seg_start[k] = s; sm.h.x.off[k] = (unsigned short)min(len, 65535u); }
        nseg += __popc(m);
        part += len;
    }
    const unsigned total = (unsigned)warp_sum((int)part);
    filled = false;
    if (total < fill_lo || total > fill_hi || nseg > (unsigned)KN_SEGCAP) return total;
    __syncwarp();
    {   // 2. exclusive scan (lane l owns a contiguous run of segments)
        const unsigned per = (nseg + 31) / 32;
        const unsigned b = lane * per, e = min(nseg, b + per);
        unsigned sum = 0;
        for (unsigned i = b; i < e; ++i) sum += sm.h.x.off[i];
        unsigned inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += up;
        }
        unsigned run = inc - sum;
        for (unsigned i = b; i < e; ++i) {
            const unsigned len = sm.h.x.off[i];
            sm.h.x.off[i] = (unsigned short)run;
            run += len;
        }
    }
    const unsigned nchunk = (total + 31) >> 5;  // <= 32
    sm.h.x.heads[lane] = 0u;
    __syncwarp();
    for (unsigned k = lane; k < nseg; k += 32) {  // 3. head masks
        const unsigned o = sm.h.x.off[k];
        BSHOT_ASSERT(o < total && (o >> 5) < 32u);
        atomicOr(&sm.h.x.heads[o >> 5], 1u << (o & 31));
    }
    __syncwarp();
    const unsigned hw = sm.h.x.heads[lane];
    unsigned pre = __popc(hw);  // -> exclusive prefix of the head counts over the chunks
    {
        unsigned inc = pre;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += up;
        }
        pre = inc - pre;
    }
    const unsigned le_mask = lt_mask | (1u << lane);
    for (unsigned c = 0; c < nchunk; ++c) {  // 4. balanced fill
        const unsigned w = __shfl_sync(0xffffffffu, hw, c);
        const unsigned p0 = __shfl_sync(0xffffffffu, pre, c);
        const unsigned j = (c << 5) + lane;
        if (j < total) {
            const unsigned k = p0 + __popc(w & le_mask) - 1u;  // candidate 0 is a head: k >= 0
            BSHOT_ASSERT(j < (unsigned)KN_CAP && k < nseg && sm.h.x.off[k] <= j);
            sm.u.idx[j] = sm.v.seg_start[k] + (j - sm.h.x.off[k]);
        }
    }
    __syncwarp();
    filled = true;
    return total;
}

// iterate all candidates of the current sphere; f(float4 point).
template <typename F>
__device__ __forceinline__ void knn_for_each(const GridParams& g, const unsigned* __restrict__ cell_start,
                                             const float4* __restrict__ sorted, const float4& q, KnnIter& it,
                                             KnnWarpSmem& sm, unsigned lane, F&& f) {
    if (it.list) {
        const unsigned total = it.total;
        // KN_DEPTH independent candidate loads in flight per lane before the first use
        for (unsigned j = lane; j < total; j += 32 * KN_DEPTH) {
            float4 p[KN_DEPTH];
#pragma unroll
            for (int u = 0; u < KN_DEPTH; ++u) {
                const unsigned ju = j + 32u * u;
                BSHOT_ASSERT(total <= (unsigned)KN_CAP);
                p[u] = __ldg(sorted + sm.u.idx[ju < total ? ju : j]);
            }
#pragma unroll
            for (int u = 0; u < KN_DEPTH; ++u)
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
// knn_selected() as the predicate.
//   1. size the sphere: candidates are enumerated first (cheap, no point is touched) and the radius is
//      tuned until their number falls into a window around 1.7 max_nn, each step predicted from the local
//      density (count ~ r^2 on surfaces); then one sweep counts the points inside the sphere and fills a
//      256-bin sqd histogram; fewer than max_nn inside -> grow and repeat
//   2. crossing bin of the histogram (re-histogrammed inside the bin while it holds > KN_LIST candidates)
//   3. one sweep: accumulate everything below the crossing bin, collect the bin; rank the short list by
//      key, accumulate its first (max_nn - below) entries (re-read from the original-order array `pts`)
template <typename Acc>
__device__ __forceinline__ KnnResult knn_select(const GridParams& g, const unsigned* __restrict__ cell_start,
                                                const float4* __restrict__ sorted, const float4* __restrict__ pts,
                                                const float4& q, float R, int max_nn, KnnWarpSmem& sm,
                                                unsigned lane, Acc&& acc) {
    KnnResult res;
    KnnIter& it = res.it;
    const float R2 = (float)((double)R * (double)R);
    float rho2 = R2;
    int n = 0;
    float rho = 2.0f * g.cell * 0.9999f;  // first probe
    int tries = 0;
    const float pad = BSHOT_KN_PAD * g.cell_yz;
    // candidate window that is worth a sweep: ~74 % of the candidates of a chord-clipped row set lie inside
    // the sphere, so 1.4 .. 2.4 max_nn candidates hold max_nn points with little excess (window chosen by measurement)
    const unsigned want_lo = (unsigned)(BSHOT_KN_WLO * (float)max(max_nn, 0)), want_hi = min((unsigned)(BSHOT_KN_WHI * (float)max(max_nn, 0)), (unsigned)KN_CAP);
    // ---- 1. size the sphere ---------------------------------------------------------------------
    for (;;) {
        const bool last = (max_nn <= 0) || !(rho < R);
        it.rho = last ? R : rho;
        rho2 = last ? R2 : __fmul_rn(it.rho, it.rho);
        it.rr = row_range(g, q.y, q.z, it.rho);
        const bool settle = last || tries >= 4;  // stop tuning: take whatever holds enough candidates
        bool filled;
        it.total = knn_expand(g, cell_start, q, it.rho, it.rr, last ? 0u : want_lo, settle ? (unsigned)KN_CAP : want_hi, sm, lane,
                              filled);
        ++tries;
#ifdef BSHOT_KNN_STATS
        if (lane == 0) { atomicAdd(&g_knn_stats[5], 1ull); atomicAdd(&g_knn_stats[6], (unsigned long long)it.rr.nrows); }
#endif
        if (!last && it.total < want_lo) {  // too few candidates (no point touched yet): grow, count ~ r^2 on surfaces
            const float f = fminf(fmaxf(sqrtf(1.7f * (float)max_nn / (float)max(it.total, 1u)), 1.08f), 4.0f);
            rho = fmaxf(1.08f * rho, (rho + pad) * f - pad);
            continue;
        }
        if (!settle && it.total > want_hi) {  // too many (dense spot): shrink
            const float f = fminf(fmaxf(sqrtf(1.7f * (float)max_nn / (float)it.total), 0.25f), 0.95f);
            rho = fminf(0.95f * rho, fmaxf(0.25f * rho, (rho + pad) * f - pad));
            continue;
        }
        it.list = filled;
        it.cached = false;
        for (unsigned b = lane; b < KN_BINS; b += 32) sm.h.hist[b] = 0;
        __syncwarp();
        const float scale = (float)KN_BINS / rho2;
        int cnt = 0;
        knn_for_each(g, cell_start, sorted, q, it, sm, lane, [&](const float4 p) {
            const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
            if (sqd < rho2) {
                ++cnt;
                atomicAdd(&sm.h.hist[min(KN_BINS - 1, (int)(sqd * scale))], 1u);
            }
        });
        n = warp_sum(cnt);
#ifdef BSHOT_KNN_STATS
        if (lane == 0) { atomicAdd(&g_knn_stats[0], 1ull); atomicAdd(&g_knn_stats[1], (unsigned long long)it.rr.nrows); atomicAdd(&g_knn_stats[2], (unsigned long long)it.total); atomicAdd(&g_knn_stats[3], (unsigned long long)n); atomicAdd(&g_knn_stats[4], it.list ? 0ull : 1ull); }
#endif
        if (last || n >= max_nn) break;
        rho *= fminf(fmaxf(sqrtf(1.3f * (float)max_nn / (float)max(n, 1)), 1.1f), 4.0f);
        tries = 4;  // grown after a sweep: no more shrinking
    }
    res.rho2 = rho2;
    res.n_in = n;
    // ---- 2. crossing bin (the level-0 histogram is already in sm.hist) ----------------------------------
    // select-all (n <= max_nn, or no cap) is the same collect sweep with every bin below `bin`
    const bool select_all = (max_nn <= 0 || n <= max_nn);
    float lo = 0.0f, hi = rho2, scale = (float)KN_BINS / rho2;
    int below = 0;  // selected-for-sure elements with sqd < lo
    int bin = KN_BINS;
    unsigned bbelow = 0;
    if (!select_all) {
        for (int iter = 0;; ++iter) {
            scale = (float)KN_BINS / (hi - lo);
            if (iter > 0) {
                for (unsigned b = lane; b < KN_BINS; b += 32) sm.h.hist[b] = 0;
                __syncwarp();
                int cb = 0;
                const float flo = lo, fhi = hi, fscale = scale;
                knn_for_each(g, cell_start, sorted, q, it, sm, lane, [&](const float4 p) {
                    const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
                    if (sqd < flo) ++cb;  // `below` is recounted exactly for the narrowed bound
                    else if (sqd < fhi) atomicAdd(&sm.h.hist[min(KN_BINS - 1, (int)((sqd - flo) * fscale))], 1u);
                });
                below = warp_sum(cb);
            }
            // locate the crossing bin: lane l owns bins 8l..8l+7
            unsigned h[8], s = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { h[k] = sm.h.hist[lane * 8 + k]; s += h[k]; }
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
            bin = __shfl_sync(0xffffffffu, found_bin, src);
            bbelow = __shfl_sync(0xffffffffu, found_below, src);
            const unsigned bcnt = __shfl_sync(0xffffffffu, found_cnt, src);
            if (bcnt <= KN_LIST || iter == 7) break;
            // narrow to the crossing bin and histogram again -- unless the bin is already only a few ulps wide (a pile
            // of equal distances cannot be split by distance: the bisection over the full key below takes over)
            const float w = (hi - lo) / (float)KN_BINS;
            const float nlo = fmaxf(lo, lo + w * (float)bin - w * 1e-3f), nhi = fminf(hi, lo + w * (float)(bin + 1) + w * 1e-3f);
            if (!(nhi - nlo > 64.0f * 1.1920929e-07f * nhi)) break;
            lo = nlo;
            hi = nhi;
        }
    }
    // ---- 3. accumulate below the bin, collect the bin, rank, accumulate the rest -----------------------
    {
        const float blo = lo, bhi = hi, bscale = scale;
        const int cbin = bin;
        if (lane == 0) sm.list_n = 0;
        __syncwarp();
        knn_for_each(g, cell_start, sorted, q, it, sm, lane, [&](const float4 p) {
            const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
            if (!(sqd < bhi)) return;
            const int b = (sqd >= blo) ? min(KN_BINS - 1, (int)((sqd - blo) * bscale)) : -1;
            if (b < cbin) acc(p);
            else if (b == cbin) {
                const unsigned slot = atomicAdd(&sm.list_n, 1u);
                if (slot < KN_LIST) sm.v.list[slot] = knn_key(sqd, p.w);
            }
        });
        __syncwarp();
    }
    if (select_all) {
        res.thr = (((unsigned long long)__float_as_uint(rho2)) << 32) - 1ull;
        res.count = n;
        return res;
    }
    const unsigned need = (unsigned)max_nn - bbelow;  // 1..bcnt
    if (sm.list_n > (unsigned)KN_LIST) {
        // More equal (or nearly equal) distances around the max_nn-th neighbour than the list holds, e.g. a pile of
        // (0,0,0) invalid returns that are all equidistant from the query: exact threshold key by bisection over the
        // crossing bin, then one sweep accumulates the `need` smallest keys of the bin.
        const float tlo = lo, thi = hi, tscale = scale;
        const int tbin = bin;
        auto in_bin = [&](float sqd) { return sqd < thi && sqd >= tlo && min(KN_BINS - 1, (int)((sqd - tlo) * tscale)) == tbin; };
        unsigned long long klo = 0ull, khi = ((unsigned long long)__float_as_uint(thi) << 32);
        while (klo < khi) {
            const unsigned long long mid = klo + ((khi - klo) >> 1);
            int cb = 0;
            knn_for_each(g, cell_start, sorted, q, it, sm, lane, [&](const float4 p) {
                const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
                if (in_bin(sqd) && knn_key(sqd, p.w) <= mid) ++cb;
            });
            cb = warp_sum(cb);
            if ((unsigned)cb >= need) khi = mid; else klo = mid + 1ull;
        }
        const unsigned long long thr = klo;
        knn_for_each(g, cell_start, sorted, q, it, sm, lane, [&](const float4 p) {
            const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
            if (in_bin(sqd) && knn_key(sqd, p.w) <= thr) acc(p);
        });
        res.thr = thr;
        res.count = max_nn;
        return res;
    }
    const unsigned ln = sm.list_n;
    if (lane == 0) sm.thr = 0ull;
    __syncwarp();
    for (unsigned e = lane; e < ln; e += 32) {
        const unsigned long long ke = sm.v.list[e];
        unsigned rank = 0;
        for (unsigned o = 0; o < ln; ++o) rank += (sm.v.list[o] < ke) ? 1u : 0u;
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

// ---- exact-order sums ------------------------------------------------------------------------------------------
// pcl::computeCentroid and pcl::computeMeanAndCovarianceMatrix add their fp32 accumulators in NEIGHBOUR ORDER
// (ascending distance).  knn_select() hands the selected points to `acc` in candidate order (the callers sum them in
// fp64: more accurate than the reference, but not bit-identical to it -- a vote next to the dividing plane may flip).
// The callers therefore materialise the selected set as sorted keys and replay the reference's sequential fp32
// additions, so seg-ratios, keypoints and normals come out bit-identical to the oracle.  Sets larger than
// KN_EXACT_CAP (max_nn > 512 or uncapped searches) keep the fp64 sums.
constexpr int KN_EXACT_CAP = 512;

struct KnnExactSmem {
    KnnWarpSmem k;
    unsigned long long skeys[KN_EXACT_CAP];
};

// keys of the selected set of `res` in ascending (sqd, index) order in skeys[0..res.count); false if it does not fit
__device__ __forceinline__ bool knn_sorted_selected(const GridParams& g, const unsigned* __restrict__ cell_start,
                                                    const float4* __restrict__ sorted, const float4& q, KnnResult& res,
                                                    KnnWarpSmem& sm, unsigned long long* skeys, unsigned lane) {
    if (res.count > KN_EXACT_CAP) return false;
    if (lane == 0) sm.list_n = 0;
    for (unsigned i = lane; i < (unsigned)KN_EXACT_CAP; i += 32) skeys[i] = ~0ull;  // padding sorts last
    __syncwarp();
    const float rho2 = res.rho2;
    const unsigned long long thr = res.thr;
    knn_for_each(g, cell_start, sorted, q, res.it, sm, lane, [&](const float4 p) {
        const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
        const unsigned long long key = knn_key(sqd, p.w);
        if (sqd < rho2 && key <= thr) {
            const unsigned slot = atomicAdd(&sm.list_n, 1u);
            if (slot < (unsigned)KN_EXACT_CAP) skeys[slot] = key;
        }
    });
    __syncwarp();
    BSHOT_ASSERT(sm.list_n == (unsigned)res.count);
    // bitonic sort of the smallest power of two that holds the set
    unsigned m = 32;
    while (m < (unsigned)res.count) m <<= 1;
    for (unsigned size = 2; size <= m; size <<= 1) {
        for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
            for (unsigned t = lane; t < (m >> 1); t += 32) {
                const unsigned lo = 2 * t - (t & (stride - 1));
                const unsigned hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = skeys[lo], b = skeys[hi];
                if ((a > b) == up) { skeys[lo] = b; skeys[hi] = a; }
            }
            __syncwarp();
        }
    }
    return true;
}

// replays `for (j in neighbour order) acc(point j)` with every lane holding the same running state: f(x, y, z)
// is called count times, in order, with the coordinates broadcast from the lane that loaded them
template <typename F>
__device__ __forceinline__ void knn_replay_in_order(const float4* __restrict__ pts, const unsigned long long* skeys,
                                                    int count, unsigned lane, F&& f) {
#pragma unroll 1
    for (int base = 0; base < count; base += 32) {
        const int j = base + (int)lane;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < count && skeys[j] != ~0ull) p = __ldg(pts + (unsigned)(skeys[j] & 0xFFFFFFFFull));
        const int m = min(32, count - base);
#pragma unroll 1
        for (int l = 0; l < m; ++l)
            f(__shfl_sync(0xffffffffu, p.x, l), __shfl_sync(0xffffffffu, p.y, l), __shfl_sync(0xffffffffu, p.z, l));
    }
}

__device__ __forceinline__ bool knn_selected(const KnnResult& r, float sqd, float w) {
    return knn_key(sqd, w) <= r.thr;  // thr < (rho2 bits << 32) by construction
}

}  // namespace bshot

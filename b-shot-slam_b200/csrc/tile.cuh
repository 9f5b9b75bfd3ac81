// tile.cuh -- block-tiled exact "nearest <= max_nn inside radius R" neighbourhoods for queries that ARE cloud points.
//
// This is pcl::KdTreeFLANN::radiusSearch(p, R, idx, sqd, max_nn) as the reference calls it for every surface point
// (src/lidar_odometry.cpp:70) and for the keypoints (include/bshot_bits.h:68), INCLUDING the
// order of the result (ascending fp32 squared distance, then point index): pcl::computeCentroid (:76) and
// pcl::computeMeanAndCovarianceMatrix add their fp32 accumulators in that order, so the order is part of the result.
//
// Work decomposition (grid.cu builds the block list):
//   * the voxel grid is cut into BLOCKS: cubes of 8^3 / 4^3 / 2^3 / 1 cells (coarsest cube that holds <= TL_QCAP points;
//     denser single cells are sliced).  One CTA per block.
//   * the CTA stages every point within rho of the block's bounding box ONCE into shared memory (the TILE; row segments
//     of the voxel table -> coalesced float4 reads, filtered against the rounded box), rho predicted from the local density.
//   * one warp per query, everything from shared memory:
//       sweep A   squared distances -> 1024-bin histogram over [0, rho^2)            (count inside the sphere)
//       scan      crossing bin b* of the max_nn-th neighbour, bin start offsets
//       sweep B   counting-sort scatter of every point in bins <= b* (sqd, tile slot)
//       rank      exact (sqd, index) order inside each bin -> neighbour list in the reference's order
//       gather    coordinates in neighbour order (SoA) -> sequential fp32 sums replayed by lanes, votes in parallel
//     a query whose sphere holds fewer than max_nn points while rho < R is retried with a larger tile.
//   * what does not fit (tile or segment list overflow) goes to a fallback list handled by the warp-per-query kernels
//     of knn.cuh.
#pragma once
#include "nbr.cuh"

namespace bshot {

constexpr int TL_WARPS = 4;      // shared tiles: 4 warps, 1024 points (6 CTAs per SM)
constexpr int TL_CAP = 1024;
#ifndef BSHOT_TL_QCAP
#define BSHOT_TL_QCAP 64
#endif
constexpr int TL_QCAP = BSHOT_TL_QCAP;  // queries per block
constexpr int TL_BINS = 1024;    // sqd histogram bins (u16 counters, two per word)
constexpr int TL_ENT = 352;      // entries the counting sort can hold (max_nn + crossing-bin overshoot)
constexpr int TL_MAXNN = 304;    // largest max_nn the tiled path handles
constexpr int TL_ROW = 308;      // SoA row pitch in floats (rows 16-byte aligned, 3 rows on disjoint banks)
constexpr int TL_SEGCAP = 768;   // row segments per tile

#ifndef BSHOT_TL_STAGE_U
#define BSHOT_TL_STAGE_U 2   // gathers in flight per lane while a block's tile is staged
#endif
#ifndef BSHOT_TL_SAFETY
#define BSHOT_TL_SAFETY 1.15f    // first radius = density prediction x this
#endif

struct TileWarp {
    union {
        struct {
            unsigned hist[TL_BINS / 2];        // packed u16 pairs: counts, then running start offsets
            float ent_sqd[TL_ENT];
            unsigned short ent_slot[TL_ENT];
        } s;
        float soa[3][TL_ROW];                  // x[], y[], z[] of the selected points in neighbour order
    } u;
    unsigned short order[TL_MAXNN + 16];       // tile slots in neighbour order
};
static_assert(sizeof(float) * 3 * TL_ROW <= sizeof(unsigned) * (TL_BINS / 2) + sizeof(float) * TL_ENT + 2 * TL_ENT, "SoA aliases the sort scratch");

struct TileQuery {          // result of tile_select(), warp-uniform
    int count;              // size of the selected set (0: unresolved)
    int n_in;               // points inside the sphere of radius sqrt(rho2)
};

__device__ __forceinline__ unsigned tl_bin(float sqd, float scale) {
    return min((unsigned)(TL_BINS - 1), (unsigned)__float2int_rz(__fmul_rn(sqd, scale)));
}

__device__ __forceinline__ bool tl_key_less(float sa, unsigned ia, float sb, unsigned ib) {
    return sa < sb || (sa == sb && ia < ib);
}

// Selects the nearest <= max_nn tile points with sqd < rho2 of query q and leaves their tile slots in w.order[] in the
// reference's neighbour order.  `complete` says the tile holds every point closer than R (rho2 == R^2): then fewer than
// max_nn points inside is a valid (select-all) answer, otherwise the query is unresolved (count = 0).
// The tile is padded with far-away sentinels up to a multiple of 128 points (S_pad): no bounds checks in the sweeps.
__device__ __forceinline__ TileQuery tile_select(const float4* __restrict__ tile, unsigned S_pad, const float4& q, float rho2,
                                                 bool complete, int max_nn, TileWarp& w, unsigned lane) {
    TileQuery res;
    res.count = 0;
    // ---- sweep A: histogram of the squared distances ----------------------------------------------------------
    {
        uint4* h4 = reinterpret_cast<uint4*>(w.u.s.hist);
#pragma unroll
        for (int k = 0; k < TL_BINS / 2 / 4 / 32; ++k) h4[k * 32 + lane] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();
    const float scale = (float)TL_BINS / rho2;
    for (unsigned j0 = lane; j0 < S_pad; j0 += 128) {
        float4 p[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) p[u] = tile[j0 + 32u * u];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float sqd = sqdist_rn(q.x, q.y, q.z, p[u].x, p[u].y, p[u].z);
            if (sqd < rho2) {
                const unsigned b = tl_bin(sqd, scale);
                atomicAdd(&w.u.s.hist[b >> 1], 1u << ((b & 1u) << 4));
            }
        }
    }
    __syncwarp();
    // ---- scan: lane owns bins [256 k + 8 lane, +8) of every 256-bin chunk k (conflict-free 16-byte accesses); stops at
    //      the chunk that holds the max_nn-th neighbour (later bins are never looked at) -----------------------------------
    unsigned want = (unsigned)max_nn, bstar = TL_BINS - 1, total = 0, carry = 0;
    bool found = false;
#pragma unroll 1
    for (int k = 0; k < TL_BINS / 256 && !found; ++k) {
        uint4* wp = reinterpret_cast<uint4*>(w.u.s.hist) + k * 32 + lane;
        const uint4 v = *wp;
        const unsigned h[8] = {v.x & 0xFFFFu, v.x >> 16, v.y & 0xFFFFu, v.y >> 16, v.z & 0xFFFFu, v.z >> 16, v.w & 0xFFFFu, v.w >> 16};
        const unsigned s = ((h[0] + h[1]) + (h[2] + h[3])) + ((h[4] + h[5]) + (h[6] + h[7]));
        unsigned inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += up;
        }
        const unsigned excl = carry + inc - s;  // exclusive prefix of this lane's first bin
        unsigned st[8];
        st[0] = excl;
#pragma unroll
        for (int i = 1; i < 8; ++i) st[i] = st[i - 1] + h[i - 1];
        *wp = make_uint4(st[0] | (st[1] << 16), st[2] | (st[3] << 16), st[4] | (st[5] << 16), st[6] | (st[7] << 16));
        const bool cross = excl < want && excl + s >= want;  // at most one lane
        const unsigned who = __ballot_sync(0xffffffffu, cross);
        if (who) {
            unsigned fb = 0, ftot = 0;
            if (cross) {
#pragma unroll
                for (int i = 7; i >= 0; --i)
                    if (st[i] < want && st[i] + h[i] >= want) { fb = (unsigned)(k * 256) + lane * 8u + (unsigned)i; ftot = st[i] + h[i]; }
            }
            const int src = __ffs(who) - 1;
            bstar = __shfl_sync(0xffffffffu, fb, src);
            total = __shfl_sync(0xffffffffu, ftot, src);
            found = true;
        }
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (!found) {  // fewer than max_nn points inside the sphere (every chunk was scanned: carry = their number)
        res.n_in = (int)carry;
        if (carry == 0u || !complete) return res;
        want = carry;
        total = carry;
    } else {
        res.n_in = (int)total;
    }
    __syncwarp();
    unsigned n_ent;
    if (total <= (unsigned)TL_ENT) {
        // ---- sweep B: counting-sort scatter of everything in bins <= b* ---------------------------------------
        for (unsigned j0 = lane; j0 < S_pad; j0 += 128) {
            float4 p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) p[u] = tile[j0 + 32u * u];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float sqd = sqdist_rn(q.x, q.y, q.z, p[u].x, p[u].y, p[u].z);
                if (sqd < rho2) {
                    const unsigned b = tl_bin(sqd, scale);
                    if (b <= bstar) {
                        const unsigned sh = (b & 1u) << 4;
                        const unsigned pos = (atomicAdd(&w.u.s.hist[b >> 1], 1u << sh) >> sh) & 0xFFFFu;
                        BSHOT_ASSERT(pos < (unsigned)TL_ENT);
                        w.u.s.ent_sqd[pos] = sqd;
                        w.u.s.ent_slot[pos] = (unsigned short)(j0 + 32u * u);
                    }
                }
            }
        }
        __syncwarp();
        n_ent = total;
        // ---- rank inside the bins: bin b now ends at hist[b], starts where bin b - 1 ends.  Bins hold ~0.5 entries on
        //      average: the first four of a bin are compared without a loop -----------------------------------------------
        const unsigned short* st16 = reinterpret_cast<const unsigned short*>(w.u.s.hist);
        for (unsigned i = lane; i < n_ent; i += 32) {
            const float si = w.u.s.ent_sqd[i];
            const unsigned sl = w.u.s.ent_slot[i];
            const unsigned b = tl_bin(si, scale);
            const unsigned s = b ? st16[b - 1] : 0u, e = st16[b];
            unsigned rank = s;
            bool tie = false;
#pragma unroll
            for (unsigned t = 0; t < 4; ++t) {
                const unsigned j = s + t;
                const float sj = (j < e) ? w.u.s.ent_sqd[j] : __int_as_float(0x7F800000);
                rank += (sj < si) ? 1u : 0u;
                tie = tie || (sj == si && j != i);
            }
            if (e - s > 4u || tie) {  // long bin or equal distances: exact (sqd, index) order
                const unsigned ii = __float_as_uint(tile[sl].w);
                rank = s;
                for (unsigned j = s; j < e; ++j) {
                    const float sj = w.u.s.ent_sqd[j];
                    if (sj < si || (sj == si && j != i && __float_as_uint(tile[w.u.s.ent_slot[j]].w) < ii)) ++rank;
                }
            }
            if (rank < want) w.order[rank] = (unsigned short)sl;
        }
    } else {
        // ---- rare: more equal-distance candidates than the scratch holds -> exact threshold key by bisection -----
        unsigned long long lo = 0ull, hi = (unsigned long long)__float_as_uint(rho2) << 32;
        while (lo < hi) {
            const unsigned long long mid = lo + ((hi - lo) >> 1);
            int c = 0;
            for (unsigned j = lane; j < S_pad; j += 32) {
                const float4 p = tile[j];
                const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
                if (sqd < rho2 && (((unsigned long long)__float_as_uint(sqd) << 32) | __float_as_uint(p.w)) <= mid) ++c;
            }
            c = warp_sum(c);
            if ((unsigned)c >= want) hi = mid; else lo = mid + 1ull;
        }
        unsigned base = 0;
        for (unsigned j0 = 0; j0 < S_pad; j0 += 32) {
            const unsigned j = j0 + lane;
            const float4 p = tile[j];
            const float sqd = sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z);
            const bool keep = sqd < rho2 && (((unsigned long long)__float_as_uint(sqd) << 32) | __float_as_uint(p.w)) <= lo;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const unsigned pos = base + __popc(m & ((1u << lane) - 1u));
                BSHOT_ASSERT(pos < (unsigned)TL_ENT);
                w.u.s.ent_sqd[pos] = sqd;
                w.u.s.ent_slot[pos] = (unsigned short)j;
            }
            base += __popc(m);
        }
        __syncwarp();
        n_ent = base;  // == want: the keys are distinct
        for (unsigned i = lane; i < n_ent; i += 32) {
            const float si = w.u.s.ent_sqd[i];
            const unsigned sl = w.u.s.ent_slot[i];
            const unsigned ii = __float_as_uint(tile[sl].w);
            unsigned rank = 0;
            for (unsigned j = 0; j < n_ent; ++j) {
                const float sj = w.u.s.ent_sqd[j];
                if (sj < si || (sj == si && __float_as_uint(tile[w.u.s.ent_slot[j]].w) < ii)) ++rank;
            }
            if (rank < want) w.order[rank] = (unsigned short)sl;
        }
    }
    __syncwarp();
    res.count = (int)want;
    return res;
}

// coordinates of the selected points in neighbour order -> w.u.soa (overwrites the sort scratch)
__device__ __forceinline__ void tile_gather(const float4* __restrict__ tile, int count, TileWarp& w, unsigned lane) {
    for (int r = (int)lane; r < count; r += 32) {
        const float4 p = tile[w.order[r]];
        w.u.soa[0][r] = p.x;
        w.u.soa[1][r] = p.y;
        w.u.soa[2][r] = p.z;
    }
    // zero padding up to the next multiple of 4 so that the replay can read whole float4s (adding +0.0f is exact)
    const int pad = (count + 3) & ~3;
    if ((int)lane < pad - count) {
        w.u.soa[0][count + lane] = 0.0f;
        w.u.soa[1][count + lane] = 0.0f;
        w.u.soa[2][count + lane] = 0.0f;
    }
    __syncwarp();
}

// sequential fp32 sum of row `row` over [0, count) in neighbour order (every lane may call with its own row)
__device__ __forceinline__ float tile_seq_sum(const TileWarp& w, int row, int count) {
    const float4* r4 = reinterpret_cast<const float4*>(w.u.soa[row]);
    float acc = 0.0f;
    const int n4 = (count + 3) >> 2;
#pragma unroll 2
    for (int i = 0; i < n4; ++i) {
        const float4 v = r4[i];
        acc = __fadd_rn(acc, v.x);
        acc = __fadd_rn(acc, v.y);
        acc = __fadd_rn(acc, v.z);
        acc = __fadd_rn(acc, v.w);
    }
    return acc;
}

// sequential fp32 sum of products rowa[i] * rowb[i] (rowb < 0: plain sum of rowa, only when PLAIN) in neighbour order
template <bool PLAIN = true>
__device__ __forceinline__ float tile_seq_sum_prod(const TileWarp& w, int rowa, int rowb, int count) {
    const float4* a4 = reinterpret_cast<const float4*>(w.u.soa[rowa]);
    const float4* b4 = reinterpret_cast<const float4*>(w.u.soa[(PLAIN && rowb < 0) ? rowa : rowb]);
    const bool plain = PLAIN && rowb < 0;
    float acc = 0.0f;
    const int n4 = (count + 3) >> 2;
#pragma unroll 2
    for (int i = 0; i < n4; ++i) {
        const float4 a = a4[i];
        float4 b = b4[i];
        if (plain) b = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
        acc = __fadd_rn(acc, __fmul_rn(a.x, b.x));
        acc = __fadd_rn(acc, __fmul_rn(a.y, b.y));
        acc = __fadd_rn(acc, __fmul_rn(a.z, b.z));
        acc = __fadd_rn(acc, __fmul_rn(a.w, b.w));
    }
    return acc;
}

// ---- block geometry + staging (CTA-wide) ------------------------------------------------------------------------
template <bool NRM, int CAP, int WARPS>
struct TileShared {
    static constexpr int kCap = CAP;          // tile capacity (points)
    static constexpr int kWarps = WARPS;
    static constexpr int kThreads = WARPS * 32;
    float4 tile[CAP];
    union {
        TileWarp w[WARPS];
        struct { unsigned seg_start[TL_SEGCAP]; unsigned seg_len[TL_SEGCAP]; } st;  // staging scratch (before the queries run)
    } u;
    float4 q_pt[TL_QCAP];             // the block's queries (x, y, z, surface index bits)
    unsigned q_pos[TL_QCAP];          // their positions in the cell-sorted array
    int q_nin[TL_QCAP];               // < 0: done ; >= 0: pending, n_in seen at the last try
    unsigned char q_grp[TL_QCAP];     // sub-block the query currently belongs to (blocks are split when their tile overflows)
    float nsum[NRM ? TL_QCAP : 1][10];  // normal sums (9) + count of the queries that want a normal
    float bbox[6];
    unsigned nq, ncand, nseg, seg_total, tile_n, pending, next_q, cur_block, ovf_slot;
    int min_nin;
    unsigned long long nbr_total;
};

// the block's own cells -> query list (with `flags`: the flagged points only; skip_origin: the detector's
// "skip the origin" rule, src/lidar_odometry.cpp:63).  desc = {ix0, iy0, iz0, cells per edge | slice << 4}
template <typename SM>
__device__ __forceinline__ void tile_block_queries(const GridParams& g, const unsigned* __restrict__ cell_start,
                                                   const float4* __restrict__ sorted, const uint4 desc, const int* __restrict__ flags,
                                                   bool skip_origin, SM& sm, unsigned tid) {
    const int L = (int)(desc.w & 15u);
    const unsigned slice = desc.w >> 4;
    if (tid == 0) { sm.nq = 0; sm.ncand = 0; }
    __syncthreads();
    // rows of the cube -> positions of its points (at most TL_QCAP by construction of the block list)
    if (tid < (unsigned)(L * L)) {
        const int iy = (int)desc.y + (int)(tid % L), iz = (int)desc.z + (int)(tid / L);
        if (iy < g.ny && iz < g.nz) {
            const unsigned row = ((unsigned)iz * g.ny + iy) * g.nx;
            const int x1 = min((int)desc.x + L, g.nx);
            unsigned s = __ldg(cell_start + row + desc.x), e = __ldg(cell_start + row + x1);
            if (L == 1) { s += slice * TL_QCAP; e = min(e, s + TL_QCAP); }
            if (e > s) {
                const unsigned base = atomicAdd(&sm.ncand, e - s);
                for (unsigned j = s; j < e && base + (j - s) < (unsigned)TL_QCAP; ++j) sm.q_pos[base + (j - s)] = j;
            }
        }
    }
    __syncthreads();
    // one thread per point: load, filter, compact
    const unsigned ncand = min(sm.ncand, (unsigned)TL_QCAP);
    unsigned pos = 0;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    bool keep = false;
    if (tid < ncand) {
        pos = sm.q_pos[tid];
        p = __ldg(sorted + pos);
        const bool origin = (p.x == 0.0f && p.y == 0.0f && p.z == 0.0f);
        keep = flags ? (__ldg(flags + pos) >= 0) : !(skip_origin && origin);
    }
    __syncthreads();
    if (keep) {
        const unsigned k = atomicAdd(&sm.nq, 1u);
        sm.q_pos[k] = pos;
        sm.q_pt[k] = p;
        sm.q_nin[k] = 0;
        sm.q_grp[k] = 0;
    }
    if (tid == 0) { sm.nbr_total = 0ull; }
    __syncthreads();
}

// the queries of an overflow record: list[start .. start + count) are positions in the cell-sorted array
template <typename SM>
__device__ __forceinline__ void tile_block_queries_list(const float4* __restrict__ sorted, const unsigned* __restrict__ list, unsigned start,
                                                        unsigned count, SM& sm, unsigned tid) {
    if (tid < count && tid < (unsigned)TL_QCAP) {
        const unsigned pos = __ldg(list + start + tid);
        sm.q_pos[tid] = pos;
        sm.q_pt[tid] = __ldg(sorted + pos);
        sm.q_nin[tid] = 0;
        sm.q_grp[tid] = 0;
    }
    if (tid == 0) { sm.nq = min(count, (unsigned)TL_QCAP); sm.nbr_total = 0ull; }
    __syncthreads();
}

// bounding box of the pending queries of sub-block `grp` (warp 0) and their number -> sm.bbox, sm.pending
template <typename SM>
__device__ __forceinline__ void tile_block_bbox(SM& sm, unsigned grp, unsigned tid) {
    const unsigned nq = sm.nq;
    if (tid < 32) {
        const float inf = __int_as_float(0x7F800000);
        float mn[3] = {inf, inf, inf}, mx[3] = {-inf, -inf, -inf};
        unsigned cnt = 0;
        for (unsigned k = tid; k < nq; k += 32) {
            if (sm.q_nin[k] < 0 || sm.q_grp[k] != grp) continue;
            const float4 p = sm.q_pt[k];
            mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
            mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
            ++cnt;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
                mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
            }
        cnt = (unsigned)warp_sum((int)cnt);
        if (tid == 0) {
            sm.bbox[0] = mn[0]; sm.bbox[1] = mn[1]; sm.bbox[2] = mn[2];
            sm.bbox[3] = mx[0]; sm.bbox[4] = mx[1]; sm.bbox[5] = mx[2];
            sm.pending = cnt;
        }
    }
    __syncthreads();
}

// splits the pending queries of sub-block `grp` at the middle of the bounding box's longest edge: the lower half keeps
// `grp`, the upper half becomes `new_grp`.  Returns false when they cannot be separated (coincident points).
template <typename SM>
__device__ __forceinline__ bool tile_block_split(SM& sm, unsigned grp, unsigned new_grp, unsigned tid) {
    const float ex = sm.bbox[3] - sm.bbox[0], ey = sm.bbox[4] - sm.bbox[1], ez = sm.bbox[5] - sm.bbox[2];
    const int axis = (ex >= ey && ex >= ez) ? 0 : (ey >= ez ? 1 : 2);
    const float ext = axis == 0 ? ex : (axis == 1 ? ey : ez);
    if (!(ext > 0.0f) || sm.pending < 2u) return false;
    const float mid = sm.bbox[axis] + 0.5f * ext;
    for (unsigned k = tid; k < sm.nq; k += SM::kThreads) {
        if (sm.q_nin[k] < 0 || sm.q_grp[k] != grp) continue;
        const float4 p = sm.q_pt[k];
        const float v = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
        if (v > mid) sm.q_grp[k] = (unsigned char)new_grp;  // the point at the box minimum stays, the one at the maximum moves
    }
    __syncthreads();
    return true;
}

// row segments of the voxel table that can hold points within `rs` of the bounding box; returns false when the
// segment list overflows.  sm.seg_total = number of candidate points.
template <typename SM>
__device__ __forceinline__ bool tile_enumerate(const GridParams& g, const unsigned* __restrict__ cell_start, float rs, SM& sm,
                                               unsigned tid) {
    if (tid == 0) { sm.nseg = 0; sm.seg_total = 0; sm.tile_n = 0; }
    __syncthreads();
    const float lox = sm.bbox[0], loy = sm.bbox[1], loz = sm.bbox[2], hix = sm.bbox[3], hiy = sm.bbox[4], hiz = sm.bbox[5];
    const int iy_lo = max(cell_coord(loy - rs, g.oy, g.inv_cell_yz), 0), iy_hi = min(cell_coord(hiy + rs, g.oy, g.inv_cell_yz), g.ny - 1);
    const int iz_lo = max(cell_coord(loz - rs, g.oz, g.inv_cell_yz), 0), iz_hi = min(cell_coord(hiz + rs, g.oz, g.inv_cell_yz), g.nz - 1);
    const int ny_span = iy_hi - iy_lo + 1, nz_span = iz_hi - iz_lo + 1;
    const int nrows = (ny_span > 0 && nz_span > 0) ? ny_span * nz_span : 0;
    const float eps = 1e-3f * g.cell_yz;
    unsigned mine = 0;
    for (int r = (int)tid; r < nrows; r += SM::kThreads) {
        const int iy = iy_lo + r % ny_span, iz = iz_lo + r / ny_span;
        const float y0 = g.oy + (float)iy * g.cell_yz, z0 = g.oz + (float)iz * g.cell_yz;
        // gap between the row's cross-section and the box, shrunk by eps (conservative)
        const float dy = fmaxf(fmaxf(y0 - hiy, loy - (y0 + g.cell_yz)) - eps, 0.0f);
        const float dz = fmaxf(fmaxf(z0 - hiz, loz - (z0 + g.cell_yz)) - eps, 0.0f);
        const float rem = rs * rs - dy * dy - dz * dz;
        if (rem < 0.0f) continue;
        const float xr = sqrtf(rem) + eps;
        const int ix_lo = max(cell_coord(lox - xr, g.ox, g.inv_cell), 0), ix_hi = min(cell_coord(hix + xr, g.ox, g.inv_cell), g.nx - 1);
        if (ix_lo > ix_hi) continue;
        const unsigned row = ((unsigned)iz * g.ny + iy) * g.nx;
        const unsigned s = __ldg(cell_start + row + ix_lo), e = __ldg(cell_start + row + ix_hi + 1);
        if (e > s) {
            const unsigned k = atomicAdd(&sm.nseg, 1u);
            if (k < (unsigned)TL_SEGCAP) { sm.u.st.seg_start[k] = s; sm.u.st.seg_len[k] = e - s; }
            mine += e - s;
        }
    }
    mine = (unsigned)warp_sum((int)mine);
    if ((tid & 31) == 0 && mine) atomicAdd(&sm.seg_total, mine);
    __syncthreads();
    return sm.nseg <= (unsigned)TL_SEGCAP;
}

// copies the enumerated candidates that lie within `rs` of the bounding box into the tile (one warp per segment) and
// pads it with far-away sentinels up to a multiple of 128; sm.tile_n = number staged (may exceed TL_CAP: overflow,
// tile unusable)
template <typename SM>
__device__ __forceinline__ void tile_stage(const float4* __restrict__ sorted, float rs, SM& sm, unsigned tid) {
    const unsigned lane = tid & 31, wid = tid >> 5;
    const float lox = sm.bbox[0], loy = sm.bbox[1], loz = sm.bbox[2], hix = sm.bbox[3], hiy = sm.bbox[4], hiz = sm.bbox[5];
    const float rs2 = rs * rs;
    const unsigned nseg = sm.nseg;
    for (unsigned k = wid; k < nseg; k += SM::kWarps) {
        const unsigned s = sm.u.st.seg_start[k], len = sm.u.st.seg_len[k];
        for (unsigned j0 = 0; j0 < len; j0 += 32 * BSHOT_TL_STAGE_U) {   // BSHOT_TL_STAGE_U gathers in flight per lane
            float4 p[BSHOT_TL_STAGE_U];
#pragma unroll
            for (int u = 0; u < BSHOT_TL_STAGE_U; ++u) {
                const unsigned j = j0 + 32u * u + lane;
                p[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < len) p[u] = __ldg(sorted + s + j);
            }
#pragma unroll
            for (int u = 0; u < BSHOT_TL_STAGE_U; ++u) {
                if (j0 + 32u * u >= len) break;
                const unsigned j = j0 + 32u * u + lane;
                bool keep = false;
                if (j < len) {
                    const float dx = fmaxf(fmaxf(lox - p[u].x, p[u].x - hix), 0.0f), dy = fmaxf(fmaxf(loy - p[u].y, p[u].y - hiy), 0.0f),
                                dz = fmaxf(fmaxf(loz - p[u].z, p[u].z - hiz), 0.0f);
                    keep = dx * dx + dy * dy + dz * dz <= rs2;
                }
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (m) {
                    unsigned base = 0;
                    if (lane == 0) base = atomicAdd(&sm.tile_n, (unsigned)__popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    const unsigned pos = base + __popc(m & ((1u << lane) - 1u));
                    if (keep && pos < (unsigned)SM::kCap) sm.tile[pos] = p[u];
                }
            }
        }
    }
    __syncthreads();
    const unsigned S = sm.tile_n;
    if (S <= (unsigned)SM::kCap) {
        const unsigned pad = (S + 127u) & ~127u;
        for (unsigned j = S + tid; j < pad; j += SM::kThreads) sm.tile[j] = make_float4(1e30f, 1e30f, 1e30f, __uint_as_float(0xFFFFFFFFu));
    }
    __syncthreads();
}

}  // namespace bshot

// tilek.cu -- the block-tiled neighbourhood kernel (tile.cuh): seg-ratio keypoint scores (SURVEY 8a row a2,
// src/lidar_odometry.cpp:61-126) and / or point normals (row a4, include/bshot_bits.h:43-94) for queries that are
// cloud points, with the reference's own arithmetic: fp32 running sums in neighbour order (pcl::computeCentroid :76,
// pcl::computeMeanAndCovarianceMatrix), so scores, keypoint indices and normals are bit-identical to the oracle.
#include "normal_math.cuh"
#include "stages.h"
#include "tile.cuh"

namespace bshot {

#ifndef BSHOT_TL_SPLIT_MIN
#define BSHOT_TL_SPLIT_MIN 8u   // a group with more pending queries than this is halved while its box is wider than the search radius
#endif
#ifndef BSHOT_TL_SPLIT_EDGE
#define BSHOT_TL_SPLIT_EDGE 1.0f   // a block's box is halved while its longest edge exceeds this x the search radius
#endif
#ifndef BSHOT_TS_WIDE
#define BSHOT_TS_WIDE 1000.0f   // overflow queries with at least this radius (mm) are processed first
#endif
#ifndef BSHOT_TL_MINBLOCKS
#define BSHOT_TL_MINBLOCKS 5
#endif

// Everything that follows the selection of one query's neighbourhood (w.order[0..count) = tile slots in neighbour
// order): coordinates into SoA, the reference's fp32 running sums replayed in that order, the score.  nsum (shared
// memory, 10 floats) receives the nine accumulators of pcl::computeMeanAndCovarianceMatrix + the count when NRM.
__device__ __forceinline__ float sum_gate_threshold(const SumGate& gate) {
    float thr = -INFINITY;
    if (gate.top_k > 0 && *gate.kth_count == gate.top_k) {
        const float kth = gate.kth_ratio[0];
        if (kth > 0.0f) thr = kth * 0.95f;
    }
    return thr;
}

template <int SR, bool SEG, int NRM>
__device__ __forceinline__ void tile_query_outputs(const float4* __restrict__ tile, const float4& q, int count, TileWarp& w, unsigned lane,
                                                   float* __restrict__ ratio, unsigned long long* __restrict__ keys, float* nsum,
                                                   float* __restrict__ rho_hint, int max_nn, float R, float* __restrict__ qsums = nullptr, float sum_thr = 0.0f) {
    const float nanf_ = __int_as_float(0x7FC00000);
    const unsigned qi = __float_as_uint(q.w);
    tile_gather(tile, count, w, lane);
    const float fn = (float)count;
    if (SEG && lane == 0) {
        // radius of this neighbourhood (distance of its last member), kept per point: a later pass over the same point
        // -- the normals of the points that become keypoints -- stages exactly this ball
        const int l = count - 1;
        rho_hint[qi] = (count >= max_nn) ? sqrtf(sqdist_rn(q.x, q.y, q.z, w.u.soa[0][l], w.u.soa[1][l], w.u.soa[2][l])) * 1.0001f + 0.01f : R;
    }
    float sx, sy, sz;
    constexpr bool GATED = NRM == 2 && SR == BSHOT_SR_CV;   // sums only where the score makes a keypoint likely (see SumGate)
    if (NRM && !GATED) {
        // the nine accumulators, one lane each (6..8 = plain sums = the centroid sums of pcl::computeCentroid)
        const int c = (int)(lane % 9u);
        const int rowa = (c < 3) ? 0 : (c < 5 ? 1 : (c == 5 ? 2 : c - 6));
        const int rowb = (c < 3) ? c : (c < 5 ? c - 2 : (c == 5 ? 2 : -1));
        const float acc = tile_seq_sum_prod(w, rowa, rowb, count);
        if (NRM == 2) {  // deferred: the sums of every point go to global memory, the normal is solved later for the keypoints only
            if (lane < 10) qsums[10 * (size_t)qi + lane] = (lane < 9) ? acc : fn;
        } else {
            if (lane < 9) nsum[lane] = acc;
            if (lane == 9) nsum[9] = fn;
        }
        sx = __shfl_sync(0xffffffffu, acc, 6);
        sy = __shfl_sync(0xffffffffu, acc, 7);
        sz = __shfl_sync(0xffffffffu, acc, 8);
    } else {
        const float acc = tile_seq_sum(w, (int)(lane % 3u), count);  // pcl::computeCentroid (:76)
        sx = __shfl_sync(0xffffffffu, acc, 0);
        sy = __shfl_sync(0xffffffffu, acc, 1);
        sz = __shfl_sync(0xffffffffu, acc, 2);
    }
    if (SEG && !(NRM == 1 && q.x == 0.0f && q.y == 0.0f && q.z == 0.0f)) {  // :63 the detector skips the origin
        const float vx = __fsub_rn(q.x, sx / fn), vy = __fsub_rn(q.y, sy / fn), vz = __fsub_rn(q.z, sz / fn);  // :79
        float seg;
        if (SR == BSHOT_SR_CV) {  // :83-97
            int pos = 0, neg = 0;
            for (int r = (int)lane; r < count; r += 32) {
                const float d = dot3_rn(vx, vy, vz, __fsub_rn(w.u.soa[0][r], q.x), __fsub_rn(w.u.soa[1][r], q.y), __fsub_rn(w.u.soa[2][r], q.z));
                if (d > 0.0f) ++pos;
                else if (d < 0.0f) ++neg;
            }
            pos = warp_sum(pos);
            neg = warp_sum(neg);
            const float fp = (float)pos, fq = (float)neg;
            seg = 1.0f - fminf(fp, fq) / fmaxf(fp, fq);
            if (pos == 0 && neg == 0) seg = nanf_;  // 0/0 like the reference
            if (GATED && seg >= sum_thr) {  // warp-uniform; the coordinates are still in w.u.soa
                const int c = (int)(lane % 6u);  // xx xy xz yy yz zz
                const int rowa = (c < 3) ? 0 : (c < 5 ? 1 : 2), rowb = (c < 3) ? c : (c < 5 ? c - 2 : 2);
                const float acc2 = tile_seq_sum_prod<false>(w, rowa, rowb, count);
                if (lane < 10) qsums[10 * (size_t)qi + lane] = (lane < 6) ? acc2 : (lane == 6 ? sx : (lane == 7 ? sy : (lane == 8 ? sz : fn)));
            }
        } else {  // CVS :98-108 / CVSN :109-119: per-neighbour terms in parallel, fp32 running sum in neighbour order
            const float ctn = sqrtf(dot3_rn(vx, vy, vz, vx, vy, vz));
            __syncwarp();
            for (int r = (int)lane; r < count; r += 32) {
                const float dx = __fsub_rn(w.u.soa[0][r], q.x), dy = __fsub_rn(w.u.soa[1][r], q.y), dz = __fsub_rn(w.u.soa[2][r], q.z);
                const float dn = sqrtf(dot3_rn(dx, dy, dz, dx, dy, dz));
                float term = 0.0f;  // the reference skips the neighbour (:103,:114); adding +0 is the same sum
                if (ctn != 0.0f && dn != 0.0f) {
                    const float d = dot3_rn(vx, vy, vz, dx, dy, dz);
                    term = (SR == BSHOT_SR_CVS) ? d : d / __fmul_rn(ctn, dn);
                }
                w.u.soa[0][r] = term;
            }
            __syncwarp();
            seg = fabsf(tile_seq_sum(w, 0, count)) / fn;
        }
        if (lane == 0) {
            ratio[qi] = seg;
            keys[qi] = isnan(seg) ? 0ull : (((unsigned long long)__float_as_uint(seg) << 32) | (unsigned)(~qi));
        }
    }
}

// ctl = {[0] heavy blocks (front of the list), [1] fallback-list length, [2] work counter (blocks), [3] work counter
//        (overflow queries), [4] overflow queries, [5] light blocks (back of the list), [6] wide overflow queries (back of ovf)}
// counters: [0] / [1] selected neighbours (detector / normals), [2] tiles staged, [3] tile points swept, [4] query
// attempts, [5] attempts that found fewer than max_nn points, [6] queries handed to the fallback, [7] blocks.

// SR: score type.  SEG: write ratio / keys for every (non-origin) point of the block.  NRM: write normals --
// for every query (flags == nullptr, out index = surface index) or for the flagged ones (flags[sorted position] =
// keypoint ordinal >= 0, out index = ordinal: the reference's placement, include/bshot_bits.h:79-81).
// One CTA per block of the grid's block list, one shared 1024-point tile, one warp per query.  Queries whose block tile
// does not fit (very dense spots, density jumps) go to the overflow list {position, radius} for tile_single_kernel.
template <int SR, bool SEG, int NRM>
__global__ void __launch_bounds__(TL_WARPS * 32, NRM == 1 ? BSHOT_TL_MINBLOCKS : BSHOT_TL_MINBLOCKS + 1)
tile_kernel(const GridParams* __restrict__ gp, const unsigned* __restrict__ cell_start, const float4* __restrict__ sorted,
            const uint4* __restrict__ blocks, const float* __restrict__ blk_area, unsigned* __restrict__ ctl, float radius, int max_nn,
            float* __restrict__ ratio, unsigned long long* __restrict__ keys, const int* __restrict__ flags, float4* __restrict__ nrm_out,
            unsigned long long* __restrict__ counters, uint2* __restrict__ ovf, unsigned block_cap, float* __restrict__ rho_hint,
            float* __restrict__ qsums, SumGate gate) {
    using SM = TileShared<NRM == 1, TL_CAP, TL_WARPS>;
    __shared__ SM sm;
    unsigned& s_block = sm.cur_block;
    unsigned& s_slot = sm.ovf_slot;
    unsigned* const work = ctl + 2;
    const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const GridParams g = *gp;
    const float R = radius;
    const float R2 = (float)((double)R * (double)R);
    const float sum_thr = sum_gate_threshold(gate);
    unsigned cum[kBlockClasses + 1];  // blocks by weight class, heaviest class first
    cum[0] = 0;
#pragma unroll
    for (int cl = 0; cl < kBlockClasses; ++cl) cum[cl + 1] = cum[cl] + ctl[16 + cl];
    const unsigned nb = cum[kBlockClasses];
    unsigned long long st_staged = 0, st_swept = 0, st_attempts = 0, st_blocks = 0;  // thread 0 only
    for (;;) {
        __syncthreads();  // previous block's epilogue is done with the shared state
        if (tid == 0) s_block = atomicAdd(work, 1u);  // blocks differ widely in cost: dynamic assignment, heavy blocks first
        __syncthreads();
        if (s_block >= nb) break;
        unsigned cls = 0;
#pragma unroll
        for (int cl = 1; cl < kBlockClasses; ++cl) cls += (s_block >= cum[cl]) ? 1u : 0u;
        unsigned base = 0;
#pragma unroll
        for (int cl = 1; cl < kBlockClasses; ++cl) base = (cls == (unsigned)cl) ? cum[cl] : base;
        const unsigned b = cls * block_cap + (s_block - base);
        const uint4 desc = __ldg(blocks + b);
        tile_block_queries(g, cell_start, sorted, desc, SEG ? nullptr : flags, SEG && NRM != 1, sm, tid);
        const unsigned nq = sm.nq;
        if (nq == 0) continue;
        ++st_blocks;
        // first radius: max_nn points on a surface of `area` mm^2 per point
        float rho = BSHOT_TL_SAFETY * sqrtf((float)max_nn * __ldg(blk_area + b) * 0.31830989f);
        const float rho_block = rho;
        // sub-blocks still to do (bit g = group g)
        unsigned long long todo = 1ull;
        unsigned ngrp = 1;
        while (todo) {
            const unsigned grp = (unsigned)__ffsll((long long)todo) - 1u;
            todo &= todo - 1ull;
#ifdef BSHOT_TL_GROUP_RHO
            rho = rho_block;  // every group starts from the block's estimate, not from what a sparser group grew to
#endif
            tile_block_bbox(sm, grp, tid);
            if (sm.pending == 0) continue;
            // a box much wider than the search radius makes every query sweep far more candidates than its own
            // neighbourhood: halve it first (the tile of each half is smaller; staging is cheap next to the sweeps)
            if (sm.pending > BSHOT_TL_SPLIT_MIN && ngrp < 64u) {
                const float edge = fmaxf(fmaxf(sm.bbox[3] - sm.bbox[0], sm.bbox[4] - sm.bbox[1]), sm.bbox[5] - sm.bbox[2]);
                if (edge > BSHOT_TL_SPLIT_EDGE * fminf(rho, R) && tile_block_split(sm, grp, ngrp, tid)) {
                    todo |= (1ull << grp) | (1ull << ngrp);
                    ++ngrp;
                    continue;
                }
            }
            bool shrunk = false;
            for (;;) {
                __syncthreads();  // everyone is done with the previous attempt's shared state
                const bool at_R = !(rho < R);
                const float rr = at_R ? R : rho;
                const float rs = rr * 1.0001f + 0.1f;           // staging radius: covers fp32 rounding of the distances
                const float rho2 = at_R ? R2 : __fmul_rn(rr, rr);
                const bool ok = tile_enumerate(g, cell_start, rs, sm, tid);
                if (ok && !at_R && (float)sm.seg_total < 1.2f * (float)max_nn) {  // cannot hold max_nn points: grow before touching a point
                    rho = rr * fminf(fmaxf(sqrtf(1.6f * (float)max_nn / (float)max(sm.seg_total, 1u)), 1.2f), 3.0f);
                    continue;
                }
                // more than 2.2 x the tile before the box filter never fits: skip the copy
                const bool hopeless = !ok || sm.seg_total > (unsigned)(2.2f * (float)SM::kCap);
                if (!hopeless) tile_stage(sorted, rs, sm, tid);
                const unsigned S = hopeless ? sm.seg_total : sm.tile_n;
                if (hopeless || S > (unsigned)SM::kCap) {
                    if (!shrunk && ok && rr > 0.25f * g.cell) {  // the prediction may simply be too generous: one smaller try
                        shrunk = true;
                        rho = rr * fminf(fmaxf(sqrtf(0.6f * (float)SM::kCap / (float)S), 0.4f), 0.9f);
                        continue;
                    }
                    // does not fit a shared tile: every pending query gets its own tile (tile_single_kernel)
                    __syncthreads();
                    // wide balls (sparse spots: many rows to walk) are queued from the back of the list and taken first
                    const bool wide = rr >= BSHOT_TS_WIDE;
                    if (tid == 0) { s_slot = atomicAdd(ctl + (wide ? 6 : 4), sm.pending); sm.next_q = 0; }
                    __syncthreads();
                    for (unsigned k = tid; k < nq; k += SM::kThreads)
                        if (sm.q_nin[k] >= 0 && sm.q_grp[k] == grp) {
                            const unsigned slot = s_slot + atomicAdd(&sm.next_q, 1u);
                            ovf[wide ? block_cap - 1u - slot : slot] = make_uint2(sm.q_pos[k], __float_as_uint(fminf(rr, R)));
                            sm.q_nin[k] = -2;
                        }
                    break;
                }
                if (!at_R && S < (unsigned)max_nn) {
                    rho = rr * fminf(fmaxf(sqrtf(1.5f * (float)max_nn / (float)max(S, 8u)), 1.2f), 3.0f);
                    continue;
                }
                const unsigned S_pad = (S + 127u) & ~127u;
                if (tid == 0) { sm.min_nin = 0x7FFFFFFF; sm.next_q = 0; }
                ++st_staged; st_swept += (unsigned long long)S * sm.pending; st_attempts += sm.pending;
                __syncthreads();
                shrunk = true;  // from here on the radius only grows
                TileWarp& w = sm.u.w[wid];
                for (;;) {
                    unsigned k = 0;
                    if (lane == 0) k = atomicAdd(&sm.next_q, 1u);  // queries differ in cost too
                    k = __shfl_sync(0xffffffffu, k, 0);
                    if (k >= nq) break;
                    if (sm.q_nin[k] < 0 || sm.q_grp[k] != grp) continue;
                    const float4 q = sm.q_pt[k];
                    const TileQuery tq = tile_select(sm.tile, S_pad, q, rho2, at_R, max_nn, w, lane);
                    if (tq.count == 0) {
                        if (lane == 0) { sm.q_nin[k] = max(tq.n_in, 1); atomicMin(&sm.min_nin, tq.n_in); atomicAdd(&counters[5], 1ull); }
                        continue;
                    }
                    tile_query_outputs<SR, SEG, NRM>(sm.tile, q, tq.count, w, lane, ratio, keys, NRM == 1 ? sm.nsum[k] : nullptr, rho_hint, max_nn, R, qsums, sum_thr);
                    if (lane == 0) {
                        sm.q_nin[k] = -1;
                        atomicSub(&sm.pending, 1u);
                        atomicAdd(&sm.nbr_total, (unsigned long long)tq.count);
                    }
                    __syncwarp();
                }
                __syncthreads();
                if (sm.pending == 0) break;
                // some spheres held fewer than max_nn points: grow by the density they saw (count ~ r^2 on surfaces)
                rho = rr * fminf(fmaxf(sqrtf(1.35f * (float)max_nn / (float)max(sm.min_nin, 1)), 1.15f), 3.0f);
#ifndef BSHOT_TL_NO_REBOX
                tile_block_bbox(sm, grp, tid);  // the box of the queries that are left: the larger radius stages far less around it
#endif
            }
        }
        __syncthreads();
        if (NRM == 1) {  // eigen-solves of the block's queries in parallel
            for (unsigned k = tid; k < nq; k += SM::kThreads) {
                if (sm.q_nin[k] != -1) continue;
                const float4 q = sm.q_pt[k];
                float a[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) a[i] = sm.nsum[k][i];
                const float4 o = normal_from_sums(a, (int)sm.nsum[k][9], q.x, q.y, q.z);
                const unsigned oi = flags ? (unsigned)__ldg(flags + sm.q_pos[k]) : __float_as_uint(q.w);
                nrm_out[oi] = o;
            }
        }
        if (tid == 0 && sm.nbr_total) atomicAdd(&counters[SEG ? 0 : 1], sm.nbr_total);
    }
    if (tid == 0) {
        atomicAdd(&counters[2], st_staged); atomicAdd(&counters[3], st_swept); atomicAdd(&counters[4], st_attempts);
        atomicAdd(&counters[7], st_blocks);
    }
}

// ---- one private tile per query ------------------------------------------------------------------------------------
// The overflow queries of tile_kernel: each warp takes ONE query at a time, stages the ball around it into its own
// TS_CAP-point tile (row segments read by the warp, no CTA-wide step) and brackets the radius until the ball holds at
// least max_nn and at most TS_CAP points -- always possible unless more than TS_CAP points coincide in distance, which
// goes to the warp-per-query fallback list (knn.cuh).
#ifndef BSHOT_TS_FLAT_BELOW
#define BSHOT_TS_FLAT_BELOW 32u   // mean points per non-empty row segment below which the segments are read as one range
#endif
#ifdef BSHOT_TS_DEBUG
__device__ long long g_ts_dbg[8 * 8192];
void ts_debug_dump() {
    static long long h[8 * 8192];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(h, g_ts_dbg, sizeof(h));
    long long t0 = 0x7fffffffffffffffll;
    for (int i = 0; i < 8192; ++i) if (h[8 * i + 4] && h[8 * i + 6] < t0) t0 = h[8 * i + 6];
    for (int i = 0; i < 8192; ++i) if (h[8 * i + 4]) fprintf(stderr, "TS %d rho0 %lld final %lld it %lld S %lld cyc %lld stage %lld start %lld warp %lld\n", i, h[8 * i] / 10, h[8 * i + 1] / 10, h[8 * i + 2], h[8 * i + 3], h[8 * i + 4], h[8 * i + 5], h[8 * i + 6] - t0, h[8 * i + 7]);
    static long long z[8 * 8192];
    cudaMemcpyToSymbol(g_ts_dbg, z, sizeof(z));
}
#endif
constexpr int TS_WARPS = 4;
constexpr int TS_CAP = 640;

struct SingleWarp {
    float4 tile[TS_CAP];
    TileWarp w;
    float nsum[12];
};

// stages every point with distance <= rs of q into st.tile; returns the number found (may exceed TS_CAP: tile unusable)
// stages the ball of radius rs around q into the warp's tile; returns the number of points inside (may exceed TS_CAP: then
// the tile is unusable).  st.w.u.s.hist[0..63] receives the histogram of the squared distances in 64 equal steps of rs^2,
// from which the caller picks a radius that fits when this one did not.
constexpr int TS_HBINS = 64;
__device__ __forceinline__ unsigned single_stage(const GridParams& g, const unsigned* __restrict__ cell_start, const float4* __restrict__ sorted,
                                                 const float4& q, float rs, SingleWarp& st, unsigned lane) {
    const RowRange rr = row_range(g, q.y, q.z, rs);
    const float rs2 = rs * rs, hscale = (float)TS_HBINS / rs2;
    unsigned* const hist = st.w.u.s.hist;
    hist[lane] = 0u; hist[lane + 32] = 0u;
    __syncwarp();
    unsigned n = 0;
    auto segment_of = [&](int r, unsigned& s_out, unsigned& len_out) {
        s_out = 0; len_out = 0;
        if (r < rr.nrows) {
            int iy, iz;
            row_coords(rr, r, iy, iz);
            unsigned e;
            if (row_segment(g, cell_start, q.x, q.y, q.z, rs, iy, iz, s_out, e)) len_out = e - s_out;
        }
    };
    auto take = [&](const float4& p, bool valid) {
        bool keep = false;
        float d2 = 0.0f;
        if (valid) {
            const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
            d2 = dx * dx + dy * dy + dz * dz;
            keep = d2 <= rs2;
        }
        const unsigned km = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const unsigned pos = n + __popc(km & ((1u << lane) - 1u));
            if (pos < (unsigned)TS_CAP) st.tile[pos] = p;
            atomicAdd(&hist[min((int)(d2 * hscale), TS_HBINS - 1)], 1u);
        }
        n += __popc(km);
    };
    unsigned s_next, len_next;
    segment_of((int)lane, s_next, len_next);
    for (int r0 = 0; r0 < rr.nrows; r0 += 32) {
        const unsigned s = s_next, len = len_next;
        if (r0 + 32 < rr.nrows) segment_of(r0 + 32 + (int)lane, s_next, len_next);  // the next batch's table reads fly during this one's gathers
        unsigned incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += up;
        }
        const unsigned excl = incl - len, total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned m = __ballot_sync(0xffffffffu, len > 0);
        if (total >= BSHOT_TS_FLAT_BELOW * (unsigned)__popc(m)) {
            // long segments (dense spots): one segment after the other, 64 points a step
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1u;
                const unsigned ss = __shfl_sync(0xffffffffu, s, src), ll = __shfl_sync(0xffffffffu, len, src);
                for (unsigned j0 = 0; j0 < ll; j0 += 128) {   // four gathers in flight per lane
                    float4 p[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const unsigned j = j0 + 32u * u + lane;
                        p[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (j < ll) p[u] = __ldg(sorted + ss + j);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (j0 + 32u * u < ll) take(p[u], j0 + 32u * u + lane < ll);
                }
            }
            continue;
        }
        // short segments: the 32 of them are read as ONE concatenated index range (a round trip to memory per 64 points,
        // not per segment) -- every lane finds the segment of its index by bisection over the exclusive prefix
        auto locate = [&](unsigned j) {
            unsigned lo = 0;  // last segment whose first index is <= j (empty segments share the index of their successor)
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const unsigned cand = lo + (unsigned)step;
                const unsigned e = __shfl_sync(0xffffffffu, excl, cand & 31u);
                if (cand < 32u && e <= j) lo = cand;
            }
            const unsigned ss = __shfl_sync(0xffffffffu, s, lo), se = __shfl_sync(0xffffffffu, excl, lo);
            return ss + (j - se);
        };
        for (unsigned j0 = 0; j0 < total; j0 += 128) {
            float4 p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned j = j0 + 32u * u + lane;
                const unsigned a = locate(j);
                p[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < total) p[u] = __ldg(sorted + a);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (j0 + 32u * u < total) take(p[u], j0 + 32u * u + lane < total);
        }
    }
    if (n <= (unsigned)TS_CAP) {
        const unsigned pad = (n + 127u) & ~127u;
        for (unsigned j = n + lane; j < pad; j += 32) st.tile[j] = make_float4(1e30f, 1e30f, 1e30f, __uint_as_float(0xFFFFFFFFu));
    }
    __syncwarp();
    return n;
}

// after a pass that found more than TS_CAP points: the largest radius (a bin edge of the pass's histogram) whose ball
// holds at most `room` points; *count_out = how many it holds.  Warp-uniform.
__device__ __forceinline__ float single_fit_radius(const SingleWarp& st, float rs, unsigned room, unsigned lane, unsigned* count_out) {
    const unsigned* hist = st.w.u.s.hist;
    const unsigned c0 = hist[2 * lane], c1 = hist[2 * lane + 1];
    unsigned incl = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += up;
    }
    const unsigned cum1 = incl, cum0 = incl - c1;   // points in bins [0, 2 lane] / [0, 2 lane + 1]
    const unsigned fit1 = __ballot_sync(0xffffffffu, cum1 <= room), fit0 = __ballot_sync(0xffffffffu, cum0 <= room);
    // cumulative counts grow with the bin: the bins that fit are a prefix
    const int nb = __popc(fit0) + __popc(fit1);   // number of leading bins whose cumulative count fits
    unsigned cnt = 0;
    if (nb > 0) {
        const int last = nb - 1;
        const unsigned v1 = __shfl_sync(0xffffffffu, cum1, last >> 1), v0 = __shfl_sync(0xffffffffu, cum0, last >> 1);
        cnt = (last & 1) ? v1 : v0;
    }
    *count_out = cnt;
    return rs * sqrtf((float)nb / (float)TS_HBINS);
}

template <int SR, bool SEG, int NRM>
__global__ void __launch_bounds__(TS_WARPS * 32, 3)
tile_single_kernel(const GridParams* __restrict__ gp, const unsigned* __restrict__ cell_start, const float4* __restrict__ sorted,
                   unsigned* __restrict__ ctl, float radius, int max_nn, float* __restrict__ ratio, unsigned long long* __restrict__ keys,
                   const int* __restrict__ flags, float4* __restrict__ nrm_out, unsigned long long* __restrict__ counters,
                   const uint2* __restrict__ ovf, unsigned* __restrict__ fb_list, float* __restrict__ rho_hint,
                   const int* __restrict__ kp_idx, const int* __restrict__ kp_count, const unsigned* __restrict__ sorted_pos,
                   float* __restrict__ qsums, SumGate gate, unsigned ovf_cap) {
    extern __shared__ __align__(16) unsigned char single_smem_raw[];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    SingleWarp& st = reinterpret_cast<SingleWarp*>(single_smem_raw)[wid];
    const GridParams g = *gp;
    const float R = radius;
    const float R2 = (float)((double)R * (double)R);
    // items: the overflow queries of tile_kernel -- or (kp_idx != nullptr) the detector's keypoints, each with the radius
    // the detector kept for it; result slot = keypoint ordinal (the reference's placement, include/bshot_bits.h:79-81)
    const float sum_thr = sum_gate_threshold(gate);
    const unsigned n_wide = kp_idx ? 0u : ctl[6];
    const unsigned n_items = kp_idx ? (unsigned)max(*kp_count, 0) : ctl[4] + n_wide;
    unsigned long long nbr = 0;
    for (;;) {
        unsigned i = 0;
        if (lane == 0) i = atomicAdd(ctl + 3, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n_items) break;
        uint2 item;
        if (kp_idx) {
            const int idx = kp_idx[i];
            item = make_uint2(sorted_pos[idx], __float_as_uint(rho_hint[idx]));
        } else {
            item = i < n_wide ? ovf[ovf_cap - 1u - i] : ovf[i - n_wide];
        }
        const float4 q = __ldg(sorted + item.x);
        float rho = __uint_as_float(item.y), rho_lo = 0.0f, rho_hi = 3.0e38f;
        bool done = false;
#ifdef BSHOT_TS_DEBUG
        int dbg_it = 0; unsigned dbg_S = 0; float dbg_rr = 0.f;
        const long long dbg_t0 = clock64(); long long dbg_stage = 0;
#endif
        for (int it = 0; it < 48 && !done; ++it) {
            const bool at_R = !(rho < R);
            const float rr = at_R ? R : rho;
            const float rs = rr * 1.0001f + 0.1f;
            const float rho2 = at_R ? R2 : __fmul_rn(rr, rr);
#ifdef BSHOT_TS_DEBUG
            const long long dbg_s0 = clock64();
#endif
            const unsigned S = single_stage(g, cell_start, sorted, q, rs, st, lane);
#ifdef BSHOT_TS_DEBUG
            dbg_it = it + 1; dbg_S = S; dbg_rr = rr; dbg_stage += clock64() - dbg_s0;
#endif
            if (S > (unsigned)TS_CAP) {  // too many: shrink
                rho_hi = fminf(rho_hi, rr);
                // the pass left the distance histogram: take the largest bin edge whose ball fits -- if that ball holds
                // max_nn points the next pass is the last one
                unsigned fit_count;
                const float fit = single_fit_radius(st, rs, (unsigned)TS_CAP - 8u, lane, &fit_count) * 0.9999f;
                if (fit_count >= (unsigned)max_nn && fit > rho_lo && fit < rho_hi) { rho = fit; continue; }
                if (fit > rho_lo && fit < rho_hi && fit_count > 0u) rho_lo = fit;   // fewer than max_nn inside: the answer lies beyond it
                if (rho_lo > 0.0f) { if (!(rho_hi > 1.0005f * rho_lo)) break; rho = sqrtf(rho_lo * rho_hi); }
                else rho = rr * fminf(fmaxf(sqrtf(0.6f * (float)TS_CAP / (float)S), 0.3f), 0.9f);
                continue;
            }
            int n_in = (int)S;
            if (at_R || S >= (unsigned)max_nn) {
                const TileQuery tq = tile_select(st.tile, (S + 127u) & ~127u, q, rho2, at_R, max_nn, st.w, lane);
                if (tq.count > 0) {
                    tile_query_outputs<SR, SEG, NRM>(st.tile, q, tq.count, st.w, lane, ratio, keys, st.nsum, rho_hint, max_nn, R, qsums, sum_thr);
                    if (NRM == 1) {
                        __syncwarp();
                        if (lane == 0) {
                            float a[9];
#pragma unroll
                            for (int k = 0; k < 9; ++k) a[k] = st.nsum[k];
                            const float4 o = normal_from_sums(a, (int)st.nsum[9], q.x, q.y, q.z);
                            nrm_out[kp_idx ? i : (flags ? (unsigned)__ldg(flags + item.x) : __float_as_uint(q.w))] = o;
                        }
                    }
                    nbr += (unsigned long long)tq.count;
                    done = true;
                    continue;
                }
                n_in = tq.n_in;
            }
            // fewer than max_nn points inside: grow by the density seen (count ~ r^2 on surfaces), below the overflowing radius
            rho_lo = fmaxf(rho_lo, rr);
            rho = rr * fminf(fmaxf(sqrtf(1.35f * (float)max_nn / (float)max(n_in, 1)), 1.1f), 3.0f);
            if (rho >= rho_hi) { if (!(rho_hi > 1.0005f * rho_lo)) break; rho = sqrtf(rho_lo * rho_hi); }
        }
#ifdef BSHOT_TS_DEBUG
        if (lane == 0 && !kp_idx && i < 8192u) {
            long long* d = g_ts_dbg + 8 * (size_t)i;
            d[0] = (long long)(__uint_as_float(item.y) * 10.f); d[1] = (long long)(dbg_rr * 10.f); d[2] = dbg_it; d[3] = dbg_S;
            d[4] = clock64() - dbg_t0; d[5] = dbg_stage; d[6] = dbg_t0; d[7] = blockIdx.x * 4 + wid;
        }
#endif
        if (!done && lane == 0) { fb_list[atomicAdd(ctl + 1, 1u)] = item.x; atomicAdd(&counters[6], 1ull); }
        __syncwarp();
    }
    if (lane == 0 && nbr) atomicAdd(&counters[SEG ? 0 : 1], nbr);
}

bool tile_path_ok(const Ctx* c, int max_nn) { return !c->force_warp_path && max_nn > 0 && max_nn <= TL_MAXNN; }

int tile_neighbourhoods(Ctx* c, int sr_type, bool seg, int nrm, float radius, int max_nn, const int* d_flags, float4* d_nrm_out, int gate_top_k) {
    const SumGate gate{c->d_kp_ratio, c->d_kp_count, gate_top_k};
    if (max_nn <= 0 || max_nn > TL_MAXNN) { set_error("tile_neighbourhoods: max_nn %d outside (0, %d]", max_nn, TL_MAXNN); return BSHOT_E_INVALID; }
    // persistent CTAs pulling blocks / queries from work counters; d_nblocks[1..4]: fallback-list length, work counters, overflow count
    BSHOT_CUDA_TRY(cudaMemsetAsync(c->d_nblocks + 1, 0, 4 * sizeof(unsigned), c->stream));
    BSHOT_CUDA_TRY(cudaMemsetAsync(c->d_nblocks + 6, 0, sizeof(unsigned), c->stream));   // [16 ..] = blocks per weight class, written by the grid build
    const size_t single_smem = sizeof(SingleWarp) * TS_WARPS;
#define BSHOT_TILE(SR, SEG, NRM)                                                                                                           \
    do {                                                                                                                                   \
        static bool attr_set = false;                                                                                                      \
        if (!attr_set) {                                                                                                                   \
            BSHOT_CUDA_TRY(cudaFuncSetAttribute(tile_single_kernel<SR, SEG, NRM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)single_smem)); \
            attr_set = true;                                                                                                               \
        }                                                                                                                                  \
        tile_kernel<SR, SEG, NRM><<<(unsigned)c->sm_count * (BSHOT_TL_MINBLOCKS + 1), TL_WARPS * 32, 0, c->stream>>>(                       \
            c->d_grid, c->d_cell_start, c->d_sorted, c->d_blocks, c->d_blk_area, c->d_nblocks, radius, max_nn, c->d_ratio, c->d_keys, d_flags, \
            d_nrm_out, c->d_counters, c->d_ovf, (unsigned)c->max_points, c->d_rho_hint, c->d_qsums, gate);                                 \
        tile_single_kernel<SR, SEG, NRM><<<(unsigned)c->sm_count * 3u, TS_WARPS * 32, single_smem, c->stream>>>(                            \
            c->d_grid, c->d_cell_start, c->d_sorted, c->d_nblocks, radius, max_nn, c->d_ratio, c->d_keys, d_flags, d_nrm_out, c->d_counters,   \
            c->d_ovf, c->d_fb_list, c->d_rho_hint, nullptr, nullptr, nullptr, c->d_qsums, gate, (unsigned)c->max_points);                  \
    } while (0)
    // nrm: 0 none, 1 normals (d_nrm_out), 2 the nine covariance sums + count of every query into c->d_qsums (detector only)
#define BSHOT_TILE_SR(SR)                                                              \
    do {                                                                               \
        if (nrm == 1) BSHOT_TILE(SR, true, 1);                                         \
        else if (nrm == 2) BSHOT_TILE(SR, true, 2);                                    \
        else BSHOT_TILE(SR, true, 0);                                                  \
    } while (0)
    if (!seg) BSHOT_TILE(BSHOT_SR_CV, false, 1);
    else if (sr_type == BSHOT_SR_CV) BSHOT_TILE_SR(BSHOT_SR_CV);
    else if (sr_type == BSHOT_SR_CVS) BSHOT_TILE_SR(BSHOT_SR_CVS);
    else BSHOT_TILE_SR(BSHOT_SR_CVSN);
#undef BSHOT_TILE_SR
#undef BSHOT_TILE
#ifdef BSHOT_TS_DEBUG
    ts_debug_dump();
#endif
    count_launch(c, 2);
    return check_launch("tile_kernel");
}

// REFERENCE-mode normals of the detector's keypoints: one private tile per keypoint, staged with the radius the detector
// kept for that point (the ball holds its max_nn neighbours and nothing else), result at the keypoint's ordinal
int tile_keypoint_normals(Ctx* c, float radius, int max_nn, float4* d_nrm_out) {
    if (max_nn <= 0 || max_nn > TL_MAXNN) { set_error("tile_keypoint_normals: max_nn %d outside (0, %d]", max_nn, TL_MAXNN); return BSHOT_E_INVALID; }
    BSHOT_CUDA_TRY(cudaMemsetAsync(c->d_nblocks + 1, 0, 4 * sizeof(unsigned), c->stream));
    const size_t single_smem = sizeof(SingleWarp) * TS_WARPS;
    static bool attr_set = false;
    if (!attr_set) {
        BSHOT_CUDA_TRY(cudaFuncSetAttribute(tile_single_kernel<BSHOT_SR_CV, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)single_smem));
        attr_set = true;
    }
    tile_single_kernel<BSHOT_SR_CV, false, 1><<<(unsigned)c->sm_count * 3u, TS_WARPS * 32, single_smem, c->stream>>>(
        c->d_grid, c->d_cell_start, c->d_sorted, c->d_nblocks, radius, max_nn, c->d_ratio, c->d_keys, nullptr, d_nrm_out, c->d_counters, c->d_ovf,
        c->d_fb_list, c->d_rho_hint, c->d_kp_idx, c->d_kp_count, c->d_sorted_pos, nullptr, SumGate{nullptr, nullptr, 0}, (unsigned)c->max_points);
    count_launch(c);
    return check_launch("tile_single_kernel (keypoints)");
}

}  // namespace bshot

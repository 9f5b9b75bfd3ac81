// hamming.cu -- brute-force Hamming correspondence search.
//
// Replaces src/lidar_odometry.cpp:212-242 + minVect (include/bshot_bits.h:6-20) of the reference.
// One distance matrix pass: every thread keeps QPT query descriptors (11 x u32 each) in registers,
// the target descriptors are streamed through shared memory by TMA bulk copies
// (cp.async.bulk + mbarrier, SASS UBLKCP) in a 4-stage ring; per (query,target) pair the kernel
// issues 11 XOR + 11 POPC and folds the distance into a packed key (distance << 23 | local index)
// so that a min/max pair keeps the top-2 with the reference's first-minimum (lowest index)
// tie-break.  Target ranges are split over blockIdx.y; a small merge kernel combines the
// per-split candidates (and, across GPUs, the all-gathered per-rank candidates).
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "stages.h"

namespace bshot {

constexpr int HM_THREADS = 256;
constexpr int HM_TILE = 128;   // targets per pipeline stage (6 KB)
constexpr int HM_STAGES = 4;
constexpr unsigned HM_IDX_BITS = 23;
// packed key = distance * HM_K + index.  HM_K is odd (not a power of two) on purpose: the multiply-adds that
// build the key then stay IMADs on the FMA pipe instead of being strength-reduced to shifts/LEAs on the ALU
// pipe, which the XOR/CSA LOP3s already saturate.  353 * HM_K < 2^32.
constexpr unsigned HM_K = (1u << HM_IDX_BITS) + 1u;
constexpr unsigned long long HM_NONE = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// COLMIN: also returns min over the thread's queries of (distance << 23 | query index) for this target
template <int QPT, bool COLMIN>
__device__ __forceinline__ uint32_t pair_update(const uint32_t (&qw)[QPT][11], const uint4 a, const uint4 b,
                                                const uint4 c, unsigned idx, uint32_t (&k1)[QPT],
                                                uint32_t (&k2)[QPT], const uint32_t (&qor)[QPT]) {
    uint32_t ck = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
        // 352-bit XOR, then a carry-save adder tree (3:2 compressors, one LOP3 each for sum and carry)
        // folds the 11 words into 1 weight-1 and 5 weight-2 words: 6 POPC instead of 11 on the
        // quarter-rate POPC pipe, the extra LOP3s go to the full-rate ALU pipe.
        const uint32_t x0 = qw[j][0] ^ a.x, x1 = qw[j][1] ^ a.y, x2 = qw[j][2] ^ a.z, x3 = qw[j][3] ^ a.w;
        const uint32_t x4 = qw[j][4] ^ b.x, x5 = qw[j][5] ^ b.y, x6 = qw[j][6] ^ b.z, x7 = qw[j][7] ^ b.w;
        const uint32_t x8 = qw[j][8] ^ c.x, x9 = qw[j][9] ^ c.y, x10 = qw[j][10] ^ c.z;
        const uint32_t s1 = xor3(x0, x1, x2), c1 = maj3(x0, x1, x2);
        const uint32_t s2 = xor3(x3, x4, x5), c2 = maj3(x3, x4, x5);
        const uint32_t s3 = xor3(x6, x7, x8), c3 = maj3(x6, x7, x8);
        const uint32_t s4 = xor3(s1, s2, s3), c4 = maj3(s1, s2, s3);
        const uint32_t s5 = xor3(s4, x9, x10), c5 = maj3(s4, x9, x10);
        // key = (popc(s5) + 2 * sum popc(c_i)) * HM_K + idx as one IMAD chain (FMA pipe)
        uint32_t key = __popc(s5) * HM_K + idx;
        key = __popc(c1) * (2u * HM_K) + key;
        key = __popc(c2) * (2u * HM_K) + key;
        key = __popc(c3) * (2u * HM_K) + key;
        key = __popc(c4) * (2u * HM_K) + key;
        key = __popc(c5) * (2u * HM_K) + key;
        const uint32_t hi = max(k1[j], key);
        k1[j] = min(k1[j], key);
        k2[j] = min(k2[j], hi);
        if (COLMIN) ck = min(ck, key + qor[j] - idx);  // same distance, query index instead of target index
    }
    return ck;
}

// grid = (query blocks, target splits). partial[(split * nq + qi) * 2 + {0,1}] = packed
// (distance << 32 | global target index), HM_NONE when the split saw fewer than 1/2 targets.
template <int QPT, bool COLMIN>
__global__ void __launch_bounds__(HM_THREADS)
hamming_top2_kernel(const uint4* __restrict__ q, unsigned nq_cap, const unsigned* __restrict__ nq_dev,
                    const uint4* __restrict__ t, unsigned nt_cap, const unsigned* __restrict__ nt_dev, unsigned chunk,
                    unsigned long long global_base, unsigned long long* __restrict__ partial, unsigned* __restrict__ gcol) {
    // nq_cap / nt_cap size the grid and the partial layout; optional device-side counts trim the work (a frame that
    // yields fewer keypoints than top_k: records beyond the count are stale and must not take part)
    const unsigned nq = nq_dev ? min(nq_cap, *nq_dev) : nq_cap;
    const unsigned nt = nt_dev ? min(nt_cap, *nt_dev) : nt_cap;
    if (blockIdx.x * (unsigned)(HM_THREADS * QPT) >= nq) return;
    __shared__ __align__(128) uint4 tile[HM_STAGES][HM_TILE * 3];
    __shared__ __align__(8) unsigned long long full[HM_STAGES];
    __shared__ unsigned col[COLMIN ? HM_STAGES : 1][COLMIN ? HM_TILE : 1];  // per-target best (distance, query) of this CTA
    if (COLMIN) {
        for (unsigned i = threadIdx.x; i < HM_STAGES * HM_TILE; i += HM_THREADS) (&col[0][0])[i] = 0xFFFFFFFFu;
    }

    const unsigned tid = threadIdx.x;
    const unsigned split = blockIdx.y;
    const unsigned t0 = split * chunk;
    const unsigned tcount = (t0 < nt) ? min(chunk, nt - t0) : 0u;
    const int ntiles = (int)((tcount + HM_TILE - 1) / HM_TILE);
    const uint4* tbase = t + (size_t)t0 * 3;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < HM_STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int it) {
        const int s = it % HM_STAGES;
        const unsigned cnt = min((unsigned)HM_TILE, tcount - (unsigned)it * HM_TILE);
        const unsigned bytes = cnt * 48u;
        mbar_expect_tx(&full[s], bytes);
        bulk_g2s(&tile[s][0], tbase + (size_t)it * HM_TILE * 3, bytes, &full[s]);
    };
    if (tid == 0) {
        for (int it = 0; it < HM_STAGES && it < ntiles; ++it) issue(it);
    }

    uint32_t qw[QPT][11];
    uint32_t k1[QPT], k2[QPT], qor[QPT];
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
        const unsigned qi = (blockIdx.x * QPT + j) * HM_THREADS + tid;
        const unsigned ql = min(qi, nq - 1);  // padding slots replicate the last real query; their results are dropped
        qor[j] = ql;
        const uint4 a = __ldg(q + (size_t)ql * 3), b = __ldg(q + (size_t)ql * 3 + 1), c = __ldg(q + (size_t)ql * 3 + 2);
        qw[j][0] = a.x; qw[j][1] = a.y; qw[j][2] = a.z; qw[j][3] = a.w;
        qw[j][4] = b.x; qw[j][5] = b.y; qw[j][6] = b.z; qw[j][7] = b.w;
        qw[j][8] = c.x; qw[j][9] = c.y; qw[j][10] = c.z;
        k1[j] = 0xFFFFFFFFu;
        k2[j] = 0xFFFFFFFFu;
    }

    for (int it = 0; it < ntiles; ++it) {
        const int s = it % HM_STAGES;
        mbar_wait(&full[s], (unsigned)(it / HM_STAGES) & 1u);
        const uint4* tp = &tile[s][0];
        const unsigned base = (unsigned)it * HM_TILE;
        const unsigned cnt = min((unsigned)HM_TILE, tcount - base);
        if (cnt == HM_TILE) {
#pragma unroll 4
            for (int tt = 0; tt < HM_TILE; ++tt) {
                const uint32_t ck = pair_update<QPT, COLMIN>(qw, tp[3 * tt], tp[3 * tt + 1], tp[3 * tt + 2], base + tt, k1, k2, qor);
                if (COLMIN) {
                    const uint32_t wm = __reduce_min_sync(0xffffffffu, ck);
                    if ((tid & 31) == 0) atomicMin(&col[s][tt], wm);
                }
            }
        } else {
            for (unsigned tt = 0; tt < cnt; ++tt) {
                const uint32_t ck = pair_update<QPT, COLMIN>(qw, tp[3 * tt], tp[3 * tt + 1], tp[3 * tt + 2], base + tt, k1, k2, qor);
                if (COLMIN) {
                    const uint32_t wm = __reduce_min_sync(0xffffffffu, ck);
                    if ((tid & 31) == 0) atomicMin(&col[s][tt], wm);
                }
            }
        }
        __syncthreads();  // every warp is done with slot s before it is refilled
        if (COLMIN && tid < cnt) {  // publish this CTA's per-target minima, reset the slot
            const unsigned v = col[s][tid];
            if (v != 0xFFFFFFFFu) atomicMin(&gcol[t0 + base + tid], v);
            col[s][tid] = 0xFFFFFFFFu;
        }
        if (tid == 0 && it + HM_STAGES < ntiles) issue(it + HM_STAGES);
    }

#pragma unroll
    for (int j = 0; j < QPT; ++j) {
        const unsigned qi = (blockIdx.x * QPT + j) * HM_THREADS + tid;
        if (qi >= nq) continue;
        unsigned long long o1 = HM_NONE, o2 = HM_NONE;
        const unsigned long long gb = global_base + t0;
        if (k1[j] != 0xFFFFFFFFu) o1 = ((unsigned long long)(k1[j] / HM_K) << 32) | (gb + (k1[j] % HM_K));
        if (k2[j] != 0xFFFFFFFFu) o2 = ((unsigned long long)(k2[j] / HM_K) << 32) | (gb + (k2[j] % HM_K));
        unsigned long long* p = partial + ((size_t)split * nq_cap + qi) * 2;
        p[0] = o1;
        p[1] = o2;
    }
}

__device__ __forceinline__ void top2_insert(unsigned long long key, unsigned long long& k1, unsigned long long& k2) {
    const unsigned long long hi = max(k1, key);
    k1 = min(k1, key);
    k2 = min(k2, hi);
}

// merge nsrc candidate pairs per query: src[(s * nq + qi) * 2 + {0,1}].  HM_MERGE_LANES lanes share a
// query (lane l takes the splits l, l + 8, ...: one round of independent 16-byte loads instead of a serial
// walk over all splits), then a butterfly merges the partial top-2 sets.  With `gcol` (fused column
// minima of a frame-sized target set) the reverse result rq = best query of the winning target is attached.
constexpr int HM_MERGE_LANES = 8;

__global__ void __launch_bounds__(256)
merge_top2_kernel(const unsigned long long* __restrict__ src, unsigned nsrc, unsigned nq, const unsigned* __restrict__ nq_dev,
                  const unsigned* __restrict__ gcol, unsigned long long global_base, bshot_cand* __restrict__ out) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned qi = t / HM_MERGE_LANES, sub = t % HM_MERGE_LANES;
    const unsigned nq_live = nq_dev ? min(nq, *nq_dev) : nq;  // rows beyond the device-side count were never written
    unsigned long long k1 = HM_NONE, k2 = HM_NONE;
    if (qi < nq_live) {
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(src);
#pragma unroll 4
        for (unsigned sp = sub; sp < nsrc; sp += HM_MERGE_LANES) {
            const ulonglong2 v = p[(size_t)sp * nq + qi];
            top2_insert(v.x, k1, k2);
            top2_insert(v.y, k1, k2);
        }
    }
#pragma unroll
    for (int o = HM_MERGE_LANES / 2; o > 0; o >>= 1) {  // whole warps run this (no early exit above)
        const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, k1, o);
        const unsigned long long o2 = __shfl_xor_sync(0xffffffffu, k2, o);
        top2_insert(o1, k1, k2);
        top2_insert(o2, k1, k2);
    }
    if (qi >= nq || sub != 0) return;
    bshot_cand c;
    c.k1 = k1; c.k2 = k2; c.rq = 0xFFFFFFFFu; c.pad = 0;
    if (gcol && k1 != HM_NONE) {
        const unsigned v = gcol[(size_t)((k1 & 0xFFFFFFFFull) - global_base)];
        if (v != 0xFFFFFFFFu) c.rq = v % HM_K;
    }
    out[qi] = c;
}

// merge per-rank candidate RECORDS (all-gather result, rank-major): keeps rq of the winning rank
__global__ void merge_cands_kernel(const bshot_cand* __restrict__ src, unsigned nranks, unsigned nq,
                                   bshot_cand* __restrict__ out) {
    const unsigned qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    unsigned long long k1 = HM_NONE, k2 = HM_NONE;
    unsigned rq = 0xFFFFFFFFu;
    for (unsigned r = 0; r < nranks; ++r) {
        const bshot_cand c = src[(size_t)r * nq + qi];
        if (c.k1 < k1) rq = c.rq;
        top2_insert(c.k1, k1, k2);
        top2_insert(c.k2, k1, k2);
    }
    bshot_cand c;
    c.k1 = k1; c.k2 = k2; c.rq = rq; c.pad = 0;
    out[qi] = c;
}

// gather the 48-byte record of every query's best target (reverse pass input)
__global__ void gather_best_kernel(const bshot_cand* __restrict__ cand, unsigned nq, const uint4* __restrict__ t,
                                   unsigned long long global_base, uint4* __restrict__ out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq * 3) return;
    const unsigned qi = i / 3, w = i % 3;
    const unsigned long long k = cand[qi].k1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (k != HM_NONE) v = __ldg(t + (size_t)((k & 0xFFFFFFFFull) - global_base) * 3 + w);
    out[i] = v;
}

__global__ void set_rq_kernel(bshot_cand* __restrict__ cand, const bshot_cand* __restrict__ rev, unsigned nq) {
    const unsigned qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    const bool has = cand[qi].k1 != HM_NONE && rev[qi].k1 != HM_NONE;
    cand[qi].rq = has ? (unsigned)(rev[qi].k1 & 0xFFFFFFFFull) : 0xFFFFFFFFu;
}

// unpack candidate records into the reference's int arrays (left_nn etc.)
__global__ void unpack_cands_kernel(const bshot_cand* __restrict__ cand, unsigned nq, int* __restrict__ idx1,
                                    int* __restrict__ d1, int* __restrict__ idx2, int* __restrict__ d2) {
    const unsigned qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    const bshot_cand c = cand[qi];
    const bool h1 = c.k1 != HM_NONE, h2 = c.k2 != HM_NONE;
    if (idx1) idx1[qi] = h1 ? (int)(c.k1 & 0xFFFFFFFFull) : -1;
    if (d1) d1[qi] = h1 ? (int)(c.k1 >> 32) : -1;
    if (idx2) idx2[qi] = h2 ? (int)(c.k2 & 0xFFFFFFFFull) : -1;
    if (d2) d2[qi] = h2 ? (int)(c.k2 >> 32) : -1;
}

// mutual-NN filter (src/lidar_odometry.cpp:234-242): ordered compaction by one CTA -- every thread owns a contiguous
// run of queries, one block-wide exclusive scan of the per-thread counts gives each run its output offset
__global__ void __launch_bounds__(1024)
mutual_pairs_kernel(const bshot_cand* __restrict__ cand, unsigned nq_cap, const unsigned* __restrict__ nq_dev,
                    int* __restrict__ pairs, int* __restrict__ count) {
    __shared__ unsigned warp_tot[32];
    const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned nq = nq_dev ? min(nq_cap, *nq_dev) : nq_cap;
    const unsigned per = (nq + blockDim.x - 1) / blockDim.x;
    const unsigned q0 = tid * per, q1 = min(nq, q0 + per);
    unsigned mine = 0;
    // up to MP_PER queries per thread (16 384 queries): the records are read once, all loads of a thread in flight together
    constexpr int MP_PER = 16;
    const bool in_regs = per <= (unsigned)MP_PER;
    unsigned long long rk[MP_PER];
    unsigned hit = 0;   // bit j: query q0 + j is a mutual pair
    if (in_regs) {
#pragma unroll
        for (int j = 0; j < MP_PER; ++j) {
            const unsigned qi = q0 + j;
            rk[j] = HM_NONE;
            unsigned rq = 0xFFFFFFFFu;
            if (qi < q1) { rk[j] = cand[qi].k1; rq = cand[qi].rq; }
            hit |= ((rk[j] != HM_NONE) && (rq == qi)) ? (1u << j) : 0u;
        }
        mine = (unsigned)__popc(hit);
    } else {
        for (unsigned qi = q0; qi < q1; ++qi) {
            const bshot_cand c = cand[qi];
            mine += ((c.k1 != HM_NONE) && (c.rq == qi)) ? 1u : 0u;
        }
    }
    unsigned inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += up;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const unsigned w = warp_tot[lane];
        unsigned winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= (unsigned)o) winc += up;
        }
        warp_tot[lane] = winc - w;
        if (lane == 31) *count = (int)winc;
    }
    __syncthreads();
    unsigned pos = warp_tot[wid] + inc - mine;
    if (in_regs) {
#pragma unroll
        for (int j = 0; j < MP_PER; ++j)
            if (hit & (1u << j)) {
                pairs[3 * pos] = (int)(q0 + j);
                pairs[3 * pos + 1] = (int)(rk[j] & 0xFFFFFFFFull);
                pairs[3 * pos + 2] = (int)(rk[j] >> 32);
                ++pos;
            }
        return;
    }
    for (unsigned qi = q0; qi < q1; ++qi) {
        const bshot_cand c = cand[qi];
        if ((c.k1 != HM_NONE) && (c.rq == qi)) {
            pairs[3 * pos] = (int)qi;
            pairs[3 * pos + 1] = (int)(c.k1 & 0xFFFFFFFFull);
            pairs[3 * pos + 2] = (int)(c.k1 >> 32);
            ++pos;
        }
    }
}

// ---- host side ----------------------------------------------------------------------------

constexpr unsigned HM_MIN_CHUNK = 32;  // smallest target range per CTA (split granularity 16)

static unsigned qblocks_for(size_t nq, int qpt) {
    const size_t per = (size_t)HM_THREADS * qpt;
    return (unsigned)((nq + per - 1) / per);
}

// queries per thread: least padding first (ties -> more queries per thread), then fewer while the
// launch could not fill the GPU (small frame-to-frame problems are spread over many small CTAs)
static int pick_qpt(size_t nq, size_t nt, int sm_count) {
    int best = 1;
    size_t best_pad = ~(size_t)0;
    const int opts[3] = {4, 2, 1};
    for (int k = 0; k < 3; ++k) {
        const size_t pad = (size_t)qblocks_for(nq, opts[k]) * HM_THREADS * opts[k];
        if (pad < best_pad) { best_pad = pad; best = opts[k]; }
    }
    const size_t max_splits = (nt + HM_MIN_CHUNK - 1) / HM_MIN_CHUNK;
    while (best > 1 && (size_t)qblocks_for(nq, best) * max_splits < (size_t)sm_count * 2) best >>= 1;
    return best;
}

// which distance-matrix kernel runs a nq x nt search: the caller's choice (bshot_set_matcher / BSHOT_MATCH_TC), otherwise the
// tensor-core pipeline once the problem fills its tiles.  Measured crossover (tools/match_sweep.py, left top-2 + reverse):
// 2 Mi pairs, or 8 Mi pairs when the POPC kernel can fuse the column minima into its single pass (nt <= 4 nq: 2048 x 2048
// 0.024 ms POPC vs 0.029 ms tensor cores, 4096 x 4096 0.050 vs 0.037 ms)
static int matcher_for(const Ctx* c, size_t nq, size_t nt) {
    if (c->match_tc >= 0) return c->match_tc;
    const unsigned long long pairs = (unsigned long long)nq * nt;
    const bool fused = nt <= 4 * nq && nt <= c->max_targets && nq <= (1u << HM_IDX_BITS);
    return (nq >= 64 && nt >= 256 && pairs >= (fused ? (1ull << 23) : (1ull << 21))) ? 2 : 0;
}

// d_q (nq records) vs d_t (nt records): top-2 candidates per query into d_out (rq untouched = none)
// the distance-matrix pass alone: per-split top-2 candidates in c->d_partial ([nsplit][nq][2]); *nsplit_out splits
static int hamming_top2_partials(Ctx* c, const void* d_q, size_t nq, const void* d_t, size_t nt, unsigned long long global_base,
                                 unsigned* d_colmin, const unsigned* d_nq, const unsigned* d_nt, unsigned* nsplit_out, size_t live_q = 0) {
    if (nq > 0xFFFFFFFFull || nt > 0xFFFFFFFFull || global_base + nt > 0x100000000ull) {
        set_error("hamming_top2: sizes exceed 32-bit index range");
        return BSHOT_E_INVALID;
    }
    // the distance matrix on the tensor cores (hamming_tc.cu); the fused column minima stay with the POPC kernel
    const int kind = d_colmin ? 0 : matcher_for(c, nq, nt);
    if (kind >= 2) return hamming_tc2_partials(c, d_q, nq, d_t, nt, global_base, d_nq, d_nt, nsplit_out, live_q);
    if (kind == 1) return hamming_tc_partials(c, d_q, nq, d_t, nt, global_base, d_nq, d_nt, nsplit_out);
    const int qpt = pick_qpt(nq, nt, c->sm_count);
    const unsigned qblocks = qblocks_for(nq, qpt);
    // whole waves: the grid is a multiple of (SMs x resident CTAs per SM) whenever the problem is big enough,
    // otherwise the last partial wave costs as much as a full one
    static int occ_cache[2][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}};
    int& occ = occ_cache[d_colmin ? 1 : 0][qpt];
    if (occ == 0) {
        cudaError_t e = cudaSuccess;
        switch (qpt * 2 + (d_colmin ? 1 : 0)) {
            case 8: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hamming_top2_kernel<4, false>, HM_THREADS, 0); break;
            case 9: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hamming_top2_kernel<4, true>, HM_THREADS, 0); break;
            case 4: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hamming_top2_kernel<2, false>, HM_THREADS, 0); break;
            case 5: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hamming_top2_kernel<2, true>, HM_THREADS, 0); break;
            case 2: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hamming_top2_kernel<1, false>, HM_THREADS, 0); break;
            default: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hamming_top2_kernel<1, true>, HM_THREADS, 0); break;
        }
        if (e != cudaSuccess || occ <= 0) { cudaGetLastError(); occ = 2; }
    }
    const size_t slots = (size_t)c->sm_count * (size_t)occ;
    size_t want = std::max<size_t>(1, (2 * slots) / qblocks);           // two waves
    if ((nt + want - 1) / want < 4 * (size_t)HM_TILE) want = std::max<size_t>(1, slots / qblocks);  // short ranges: one wave
    const size_t cap_splits = c->partial_cap / (nq * 2);
    if (cap_splits == 0) {
        set_error("hamming_top2: partial buffer too small for %zu queries", nq);
        return BSHOT_E_CAPACITY;
    }
    if (want > cap_splits) want = cap_splits;
    if (want > 65535) want = 65535;
    size_t chunk = (nt + want - 1) / want;
    chunk = (chunk + 15) / 16 * 16;
    if (chunk < HM_MIN_CHUNK) chunk = HM_MIN_CHUNK;
    if (chunk > (1u << HM_IDX_BITS)) {
        set_error("hamming_top2: %zu targets per split exceed the packed index range", chunk);
        return BSHOT_E_CAPACITY;
    }
    unsigned nsplit = (unsigned)((nt + chunk - 1) / chunk);
    if (nsplit == 0) nsplit = 1;
    dim3 grid(qblocks, nsplit);
    const uint4* q4 = reinterpret_cast<const uint4*>(d_q);
    const uint4* t4 = reinterpret_cast<const uint4*>(d_t);
    const bool colmin = d_colmin != nullptr;
    if (colmin) {
        if (nq > (1u << HM_IDX_BITS)) { set_error("hamming_top2: fused reverse pass needs nq <= 2^23"); return BSHOT_E_CAPACITY; }
        BSHOT_CUDA_TRY(cudaMemsetAsync(d_colmin, 0xFF, sizeof(unsigned) * nt, c->stream));
    }
#define BSHOT_LAUNCH_TOP2(QPT, CM)                                                                                    \
    hamming_top2_kernel<QPT, CM><<<grid, HM_THREADS, 0, c->stream>>>(q4, (unsigned)nq, d_nq, t4, (unsigned)nt, d_nt, (unsigned)chunk, \
                                                                     global_base, c->d_partial, d_colmin)
    switch (qpt) {
        case 4: if (colmin) BSHOT_LAUNCH_TOP2(4, true); else BSHOT_LAUNCH_TOP2(4, false); break;
        case 2: if (colmin) BSHOT_LAUNCH_TOP2(2, true); else BSHOT_LAUNCH_TOP2(2, false); break;
        default: if (colmin) BSHOT_LAUNCH_TOP2(1, true); else BSHOT_LAUNCH_TOP2(1, false); break;
    }
#undef BSHOT_LAUNCH_TOP2
    count_launch(c);
    *nsplit_out = nsplit;
    return check_launch("hamming_top2_kernel");
}

int hamming_top2(Ctx* c, const void* d_q, size_t nq, const void* d_t, size_t nt, unsigned long long global_base,
                 bshot_cand* d_out, unsigned* d_colmin, const unsigned* d_nq, const unsigned* d_nt) {
    if (nq == 0) return BSHOT_OK;
    unsigned nsplit = 1;
    BSHOT_TRY(hamming_top2_partials(c, d_q, nq, d_t, nt, global_base, d_colmin, d_nq, d_nt, &nsplit));
    merge_top2_kernel<<<(unsigned)((nq * HM_MERGE_LANES + 255) / 256), 256, 0, c->stream>>>(
        c->d_partial, nsplit, (unsigned)nq, d_nq, d_colmin, global_base, d_out);
    count_launch(c);
    return check_launch("merge_top2_kernel");
}

// fills d_cand[i].rq = best query (index into d_q) for the target d_cand[i].k1
int hamming_reverse(Ctx* c, const void* d_q, size_t nq, const void* d_t, unsigned long long global_base,
                    bshot_cand* d_cand, const unsigned* d_nq) {
    if (nq == 0) return BSHOT_OK;
    gather_best_kernel<<<(unsigned)((nq * 3 + 255) / 256), 256, 0, c->stream>>>(
        d_cand, (unsigned)nq, reinterpret_cast<const uint4*>(d_t), global_base, reinterpret_cast<uint4*>(c->d_gather));
    count_launch(c);
    BSHOT_TRY(check_launch("gather_best_kernel"));
    BSHOT_TRY(hamming_top2(c, c->d_gather, nq, d_q, nq, 0, c->d_cand2, nullptr, d_nq, d_nq));  // only the live queries search / are searched
    set_rq_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, c->stream>>>(d_cand, c->d_cand2, (unsigned)nq);
    count_launch(c);
    return check_launch("set_rq_kernel");
}

int hamming_merge_cands(Ctx* c, const void* d_cands, size_t nranks, size_t nq, void* d_out) {
    if (nq == 0) return BSHOT_OK;
    merge_cands_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, c->stream>>>(
        reinterpret_cast<const bshot_cand*>(d_cands), (unsigned)nranks, (unsigned)nq, reinterpret_cast<bshot_cand*>(d_out));
    count_launch(c);
    return check_launch("merge_cands_kernel");
}

int hamming_unpack(Ctx* c, const bshot_cand* d_cand, size_t nq, int* idx1, int* d1, int* idx2, int* d2) {
    if (nq == 0) return BSHOT_OK;
    unpack_cands_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, c->stream>>>(d_cand, (unsigned)nq, idx1, d1, idx2, d2);
    count_launch(c);
    return check_launch("unpack_cands_kernel");
}

int hamming_mutual_pairs(Ctx* c, const bshot_cand* d_cand, size_t nq, int* d_pairs3, int* d_count, const unsigned* d_nq) {
    mutual_pairs_kernel<<<1, 1024, 0, c->stream>>>(d_cand, (unsigned)nq, d_nq, d_pairs3, d_count);
    count_launch(c);
    return check_launch("mutual_pairs_kernel");
}

// ---- POPC pipe microbenchmark (roofline denominator for the matcher) -----------------
__global__ void popc_peak_kernel(unsigned* out, int iters) {
    unsigned x0 = threadIdx.x * 2654435761u + blockIdx.x, x1 = x0 ^ 0x9E3779B9u, x2 = x0 + 0x7F4A7C15u,
             x3 = x0 * 3u + 1u, x4 = ~x0, x5 = x0 ^ 0xDEADBEEFu, x6 = x0 + 12345u, x7 = x0 ^ 0x55555555u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = __popc(x0) | 0x10000u; x1 = __popc(x1) | 0x20000u; x2 = __popc(x2) | 0x40000u; x3 = __popc(x3) | 0x80000u;
            x4 = __popc(x4) | 0x100000u; x5 = __popc(x5) | 0x200000u; x6 = __popc(x6) | 0x400000u; x7 = __popc(x7) | 0x800000u;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int popc_peak(Ctx* c, double* out) {
    const int blocks = c->sm_count * 8, threads = 256, iters = 2048;
    unsigned* d = nullptr;
    BSHOT_CUDA_TRY(cudaMalloc(&d, sizeof(unsigned) * blocks * threads));
    cudaEvent_t e0, e1;
    BSHOT_CUDA_TRY(cudaEventCreate(&e0));
    BSHOT_CUDA_TRY(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        BSHOT_CUDA_TRY(cudaEventRecord(e0, c->stream));
        popc_peak_kernel<<<blocks, threads, 0, c->stream>>>(d, iters);
        count_launch(c);
        BSHOT_CUDA_TRY(cudaEventRecord(e1, c->stream));
        BSHOT_CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        BSHOT_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *out = (double)blocks * threads * iters * 64.0 / (best * 1e-3);
    return BSHOT_OK;
}

}  // namespace bshot

namespace bshot {
// left top-2 AND the reverse best query in one call. When the target set is not much larger than the
// query set the column minima are tracked inside the single distance-matrix pass (Q x T pairs instead
// of Q x T + Q x Q); for map-sized target sets the Q x Q reverse pass is negligible and kept separate.
int hamming_match_rq(Ctx* c, const void* d_q, size_t nq, const void* d_t, size_t nt, unsigned long long global_base,
                     bshot_cand* d_out, const unsigned* d_nq, const unsigned* d_nt) {
    if (nq == 0) return BSHOT_OK;
    if (matcher_for(c, nq, nt) == 0 && nt <= 4 * nq && nt <= c->max_targets && nq <= (1u << HM_IDX_BITS))
        return hamming_top2(c, d_q, nq, d_t, nt, global_base, d_out, reinterpret_cast<unsigned*>(c->d_right), d_nq, d_nt);
    BSHOT_TRY(hamming_top2(c, d_q, nq, d_t, nt, global_base, d_out, nullptr, d_nq, d_nt));
    return hamming_reverse(c, d_q, nq, d_t, global_base, d_out, d_nq);
}
}  // namespace bshot

namespace bshot {

// Post-merge reverse pass of the sharded search: select the queries whose merged winner lies in this
// rank's shard [global_base, global_base + nt), gather those target records, search them against all
// queries and scatter the best query index to rq_out (0xFFFFFFFF elsewhere).  An all-reduce(MIN) over
// the ranks' rq_out arrays then gives every rank the full reverse result with Q * Q / ranks pairs of
// work per rank instead of Q * Q.
__global__ void select_owned_kernel(const bshot_cand* __restrict__ merged, unsigned nq, unsigned long long lo,
                                    unsigned long long hi, const uint4* __restrict__ t, uint4* __restrict__ gathered,
                                    unsigned* __restrict__ owner_q, unsigned* __restrict__ count, unsigned* __restrict__ rq_out) {
    const unsigned qi = blockIdx.x * blockDim.x + threadIdx.x;
    bool mine = false;
    unsigned long long idx = 0;
    if (qi < nq) {
        if (rq_out) rq_out[qi] = 0xFFFFFFFFu;  // null in push mode: the owners write straight into every rank's array
        const unsigned long long k = merged[qi].k1;
        idx = k & 0xFFFFFFFFull;
        mine = (k != HM_NONE) && idx >= lo && idx < hi;
    }
    const unsigned m = __ballot_sync(0xffffffffu, mine);
    if (m == 0) return;
    const unsigned lane = threadIdx.x & 31;
    unsigned base = 0;
    const int leader = __ffs(m) - 1;
    if ((int)lane == leader) base = atomicAdd(count, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (mine) {
        const unsigned slot = base + __popc(m & ((1u << lane) - 1));
        owner_q[slot] = qi;
        const uint4* src = t + (size_t)(idx - lo) * 3;
        gathered[(size_t)slot * 3] = __ldg(src);
        gathered[(size_t)slot * 3 + 1] = __ldg(src + 1);
        gathered[(size_t)slot * 3 + 2] = __ldg(src + 2);
    }
}

__global__ void scatter_rq_kernel(const bshot_cand* __restrict__ rev, const unsigned* __restrict__ owner_q,
                                  const unsigned* __restrict__ count, unsigned* __restrict__ rq_out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count) return;
    const unsigned long long k = rev[i].k1;
    rq_out[owner_q[i]] = (k != HM_NONE) ? (unsigned)(k & 0xFFFFFFFFull) : 0xFFFFFFFFu;
}

// ---- peer-memory exchange (symmetric buffers over NVLink, one barrier instead of a collective) ---------
// every rank stores its nq candidate records into slot `rank` of EVERY rank's gather buffer
__global__ void push_cands_kernel(const bshot_cand* __restrict__ src, unsigned nq, bshot_cand* const* __restrict__ peers,
                                  unsigned nranks, unsigned rank) {
    const unsigned qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    const bshot_cand c = src[qi];
    for (unsigned p = 0; p < nranks; ++p) {
        bshot_cand* dst = peers[(p + rank) % nranks] + (size_t)rank * nq + qi;  // start at the own buffer, spread the links
        *dst = c;
    }
}

// reverse result of the winners this rank owns -> rq[query] in every rank's array (one owner per query)
__global__ void push_rq_kernel(const bshot_cand* __restrict__ rev, const unsigned* __restrict__ owner_q,
                               const unsigned* __restrict__ count, unsigned* const* __restrict__ peers, unsigned nranks,
                               unsigned rank) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count) return;
    const unsigned long long k = rev[i].k1;
    const unsigned v = (k != HM_NONE) ? (unsigned)(k & 0xFFFFFFFFull) : 0xFFFFFFFFu;
    const unsigned q = owner_q[i];
    for (unsigned p = 0; p < nranks; ++p) peers[(p + rank) % nranks][q] = v;
}

// Cross-rank barrier over a symmetric flag array (peers[r] = rank r's array of nranks words, zero at start):
// thread p tells rank p "rank `rank` has reached barrier #epoch" and waits until rank p has said the same here.
// Release / acquire at system scope order the peer stores of the kernels before it against the kernels after it on
// every rank.  Epochs only grow, so a peer that is already one barrier ahead still satisfies the wait.  The spin is
// bounded (about 2 s): a rank that never arrives raises `*timeout_flag` instead of hanging the GPU.
__global__ void peer_barrier_kernel(unsigned* const* __restrict__ peers, unsigned nranks, unsigned rank, unsigned epoch,
                                    unsigned* __restrict__ timeout_flag) {
    const unsigned p = threadIdx.x;
    if (p >= nranks) return;
    __threadfence_system();
    unsigned* theirs = peers[p] + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
    const unsigned* mine = peers[rank] + p;
    const long long t0 = clock64();
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int)(v - epoch) >= 0) break;
        if (clock64() - t0 > 4000000000ll) { *timeout_flag = epoch; break; }
    }
}

__global__ void apply_rq_kernel(bshot_cand* __restrict__ cand, const unsigned* __restrict__ rq, unsigned nq) {
    const unsigned qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi < nq) cand[qi].rq = (cand[qi].k1 != HM_NONE) ? rq[qi] : 0xFFFFFFFFu;
}

int hamming_peer_barrier(Ctx* c, const void* d_peer_flags, unsigned nranks, unsigned rank) {
    if (nranks > 32) { set_error("peer barrier: at most 32 ranks"); return BSHOT_E_INVALID; }
    ++c->peer_epoch;
    peer_barrier_kernel<<<1, 32, 0, c->stream>>>(reinterpret_cast<unsigned* const*>(d_peer_flags), nranks, rank, c->peer_epoch,
                                                 reinterpret_cast<unsigned*>(c->d_pair_count) + 3);
    count_launch(c);
    return check_launch("peer_barrier_kernel");
}

int hamming_push_cands(Ctx* c, const bshot_cand* d_cands, size_t nq, const void* d_peer_ptrs, unsigned nranks, unsigned rank) {
    if (nq == 0) return BSHOT_OK;
    push_cands_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, c->stream>>>(d_cands, (unsigned)nq,
                                                                          reinterpret_cast<bshot_cand* const*>(d_peer_ptrs), nranks, rank);
    count_launch(c);
    return check_launch("push_cands_kernel");
}

// d_rq_out: local array (then combined by the caller's all-reduce) -- or, with d_peer_rq (device array of nranks
// pointers), the owned results are stored into every rank's array directly and d_rq_out is not touched
int hamming_reverse_owned(Ctx* c, const void* d_q, size_t nq, const void* d_t, size_t nt, unsigned long long global_base,
                          const bshot_cand* d_merged, unsigned* d_rq_out, const void* d_peer_rq, unsigned nranks, unsigned rank) {
    if (nq == 0) return BSHOT_OK;
    if (nq > c->max_kp) { set_error("hamming_reverse_owned: %zu queries > capacity %zu", nq, c->max_kp); return BSHOT_E_CAPACITY; }
    unsigned* owner_q = reinterpret_cast<unsigned*>(c->d_left);          // max_kp x 4 ints: [0,nq) owner list
    unsigned* count = reinterpret_cast<unsigned*>(c->d_pair_count) + 1;
    BSHOT_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(unsigned), c->stream));
    select_owned_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, c->stream>>>(
        d_merged, (unsigned)nq, global_base, global_base + nt, reinterpret_cast<const uint4*>(d_t),
        reinterpret_cast<uint4*>(c->d_gather), owner_q, count, d_peer_rq ? nullptr : d_rq_out);
    count_launch(c);
    BSHOT_TRY(check_launch("select_owned_kernel"));
    // gathered targets act as queries, the original queries as targets; the device-side count trims the grid
    BSHOT_TRY(hamming_top2(c, c->d_gather, nq, d_q, nq, 0, c->d_cand2, nullptr, count));
    if (d_peer_rq)
        push_rq_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, c->stream>>>(c->d_cand2, owner_q, count,
                                                                           reinterpret_cast<unsigned* const*>(d_peer_rq), nranks, rank);
    else
        scatter_rq_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, c->stream>>>(c->d_cand2, owner_q, count, d_rq_out);
    count_launch(c);
    return check_launch("scatter_rq_kernel");
}

int hamming_apply_rq(Ctx* c, bshot_cand* d_cand, const unsigned* d_rq, size_t nq) {
    if (nq == 0) return BSHOT_OK;
    apply_rq_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, c->stream>>>(d_cand, d_rq, (unsigned)nq);
    count_launch(c);
    return check_launch("apply_rq_kernel");
}


// ---- fused exchange over peer memory: the sharded call in six launches --------------------------------------------
//   1 hamming_top2_kernel        shard search (per-split candidates)
//   2 merge_push_kernel          merge the splits; STORE every record into slot `rank` of every rank's gather buffer;
//                                the last CTA to finish releases flag[rank] = epoch on every rank
//   3 wait_merge_select_kernel   acquire all flags; merge the ranks' records by (distance, global index); winners that
//                                live in this rank's shard are appended to the owned list and their records gathered
//   4 hamming_top2_kernel        owned winners against all queries (grid trimmed by the device-side count)
//   5 merge_push_rq_kernel       best query of every owned winner -> rq[query] on every rank; second flag set released
//   6 wait_apply_rq_kernel       acquire the second flags; records completed
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// every CTA fences its peer stores and takes a ticket; the last one tells every rank "rank `rank` is done with epoch"
__device__ __forceinline__ void signal_when_grid_done(unsigned* ticket, unsigned* const* peer_flags, unsigned flag_off, unsigned nranks,
                                                      unsigned rank, unsigned epoch) {
    __shared__ unsigned s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if (threadIdx.x < nranks) st_release_sys(peer_flags[threadIdx.x] + flag_off + rank, epoch);
    if (threadIdx.x == 0) *ticket = 0u;
}

// thread 0 of the CTA waits until every rank has released `epoch` into flags[flag_off + r]; bounded (~2 s)
__device__ __forceinline__ void wait_all_ranks(const unsigned* flags, unsigned flag_off, unsigned nranks, unsigned epoch, unsigned* timeout_flag) {
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (unsigned r = 0; r < nranks; ++r) {
            for (;;) {
                if ((int)(ld_acquire_sys(flags + flag_off + r) - epoch) >= 0) break;
                if (clock64() - t0 > 4000000000ll) { *timeout_flag = epoch; break; }
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
merge_push_kernel(const unsigned long long* __restrict__ src, unsigned nsrc, unsigned nq, bshot_cand* __restrict__ local_out,
                  bshot_cand* const* __restrict__ peer_gather, unsigned* const* __restrict__ peer_flags, unsigned nranks, unsigned rank,
                  unsigned epoch, unsigned* __restrict__ ticket, unsigned* __restrict__ owned_count) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned qi = t / HM_MERGE_LANES, sub = t % HM_MERGE_LANES;
    unsigned long long k1 = HM_NONE, k2 = HM_NONE;
    if (qi < nq) {
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(src);
#pragma unroll 4
        for (unsigned sp = sub; sp < nsrc; sp += HM_MERGE_LANES) {
            const ulonglong2 v = p[(size_t)sp * nq + qi];
            top2_insert(v.x, k1, k2);
            top2_insert(v.y, k1, k2);
        }
    }
#pragma unroll
    for (int o = HM_MERGE_LANES / 2; o > 0; o >>= 1) {
        const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, k1, o);
        const unsigned long long o2 = __shfl_xor_sync(0xffffffffu, k2, o);
        top2_insert(o1, k1, k2);
        top2_insert(o2, k1, k2);
    }
    if (qi < nq && sub == 0) {
        bshot_cand c;
        c.k1 = k1; c.k2 = k2; c.rq = 0xFFFFFFFFu; c.pad = 0;
        local_out[qi] = c;
        for (unsigned p = 0; p < nranks; ++p) peer_gather[(p + rank) % nranks][(size_t)rank * nq + qi] = c;  // own buffer first, links spread
    }
    if (t == 0) *owned_count = 0u;  // for step 3 of this call
    signal_when_grid_done(ticket, peer_flags, 0u, nranks, rank, epoch);
}

__global__ void __launch_bounds__(256)
wait_merge_select_kernel(const bshot_cand* __restrict__ gather, const unsigned* __restrict__ flags, unsigned nranks, unsigned nq, unsigned epoch,
                         unsigned long long lo, unsigned long long hi, const uint4* __restrict__ t, bshot_cand* __restrict__ merged,
                         uint4* __restrict__ gathered, unsigned* __restrict__ owner_q, unsigned* __restrict__ count,
                         unsigned* __restrict__ timeout_flag) {
    wait_all_ranks(flags, 0u, nranks, epoch, timeout_flag);
    const unsigned qi = blockIdx.x * blockDim.x + threadIdx.x;
    bool mine = false;
    unsigned long long idx = 0;
    if (qi < nq) {
        unsigned long long k1 = HM_NONE, k2 = HM_NONE;
        for (unsigned r = 0; r < nranks; ++r) {
            const bshot_cand c = gather[(size_t)r * nq + qi];  // written by rank r's stores, ordered by the flag acquire above
            top2_insert(c.k1, k1, k2);
            top2_insert(c.k2, k1, k2);
        }
        bshot_cand c;
        c.k1 = k1; c.k2 = k2; c.rq = 0xFFFFFFFFu; c.pad = 0;
        merged[qi] = c;
        idx = k1 & 0xFFFFFFFFull;
        mine = (k1 != HM_NONE) && idx >= lo && idx < hi;
    }
    const unsigned m = __ballot_sync(0xffffffffu, mine);
    if (m == 0) return;
    const unsigned lane = threadIdx.x & 31;
    unsigned base = 0;
    const int leader = __ffs(m) - 1;
    if ((int)lane == leader) base = atomicAdd(count, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (mine) {
        const unsigned slot = base + __popc(m & ((1u << lane) - 1));
        owner_q[slot] = qi;
        const uint4* src = t + (size_t)(idx - lo) * 3;
        gathered[(size_t)slot * 3] = __ldg(src);
        gathered[(size_t)slot * 3 + 1] = __ldg(src + 1);
        gathered[(size_t)slot * 3 + 2] = __ldg(src + 2);
    }
}

__global__ void __launch_bounds__(256)
merge_push_rq_kernel(const unsigned long long* __restrict__ src, unsigned nsrc, unsigned nq_cap, const unsigned* __restrict__ count,
                     const unsigned* __restrict__ owner_q, unsigned* const* __restrict__ peer_rq, unsigned* const* __restrict__ peer_flags,
                     unsigned nranks, unsigned rank, unsigned epoch, unsigned* __restrict__ ticket) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned i = t / HM_MERGE_LANES, sub = t % HM_MERGE_LANES;
    const unsigned n = min(*count, nq_cap);
    unsigned long long k1 = HM_NONE, k2 = HM_NONE;
    if (i < n) {
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(src);
        for (unsigned sp = sub; sp < nsrc; sp += HM_MERGE_LANES) {
            const ulonglong2 v = p[(size_t)sp * nq_cap + i];
            top2_insert(v.x, k1, k2);
            top2_insert(v.y, k1, k2);
        }
    }
#pragma unroll
    for (int o = HM_MERGE_LANES / 2; o > 0; o >>= 1) {
        const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, k1, o);
        const unsigned long long o2 = __shfl_xor_sync(0xffffffffu, k2, o);
        top2_insert(o1, k1, k2);
        top2_insert(o2, k1, k2);
    }
    if (i < n && sub == 0) {
        const unsigned v = (k1 != HM_NONE) ? (unsigned)(k1 & 0xFFFFFFFFull) : 0xFFFFFFFFu;
        const unsigned q = owner_q[i];
        for (unsigned p = 0; p < nranks; ++p) peer_rq[(p + rank) % nranks][q] = v;  // exactly one owner per query
    }
    signal_when_grid_done(ticket, peer_flags, 32u, nranks, rank, epoch);
}

__global__ void __launch_bounds__(256)
wait_apply_rq_kernel(bshot_cand* __restrict__ merged, const unsigned* __restrict__ rq, unsigned nq, const unsigned* __restrict__ flags,
                     unsigned nranks, unsigned epoch, unsigned* __restrict__ timeout_flag) {
    wait_all_ranks(flags, 32u, nranks, epoch, timeout_flag);
    const unsigned qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi < nq) merged[qi].rq = (merged[qi].k1 != HM_NONE) ? __ldcg(rq + qi) : 0xFFFFFFFFu;
}

// With lazy module loading (the CUDA 12 default) the first launch of a kernel may have to wait for running kernels; a
// consumer kernel spinning on a flag that a not-yet-loaded producer kernel of ANOTHER rank in the same process must release
// would then never see it.  Loading every kernel of the sharded call up front removes that dependency.
int hamming_preload_sharded() {
    cudaFuncAttributes a;
    BSHOT_CUDA_TRY(cudaFuncGetAttributes(&a, hamming_top2_kernel<4, false>));
    BSHOT_CUDA_TRY(cudaFuncGetAttributes(&a, hamming_top2_kernel<2, false>));
    BSHOT_CUDA_TRY(cudaFuncGetAttributes(&a, hamming_top2_kernel<1, false>));
    BSHOT_CUDA_TRY(cudaFuncGetAttributes(&a, merge_push_kernel));
    BSHOT_CUDA_TRY(cudaFuncGetAttributes(&a, wait_merge_select_kernel));
    BSHOT_CUDA_TRY(cudaFuncGetAttributes(&a, merge_push_rq_kernel));
    BSHOT_CUDA_TRY(cudaFuncGetAttributes(&a, wait_apply_rq_kernel));
    return hamming_tc2_preload();
}

// region of one rank: [gather: nranks x max_q records][rq: max_q u32][flags: 64 u32]
static size_t comm_gather_bytes(const Comm& m) { return sizeof(bshot_cand) * (size_t)m.nranks * m.max_q; }
size_t comm_region_bytes(const Comm& m) { return comm_gather_bytes(m) + sizeof(unsigned) * m.max_q + sizeof(unsigned) * 64; }

int comm_set_peers(Ctx* c, void* const* region_ptrs) {
    Comm& m = c->comm;
    std::vector<void*> g(m.nranks), r(m.nranks), f(m.nranks);
    for (int p = 0; p < m.nranks; ++p) {
        unsigned char* base = reinterpret_cast<unsigned char*>(p == m.rank ? (void*)m.d_region : region_ptrs[p]);
        if (!base) { set_error("bshot_comm: null region for rank %d", p); return BSHOT_E_INVALID; }
        g[p] = base;
        r[p] = base + comm_gather_bytes(m);
        f[p] = base + comm_gather_bytes(m) + sizeof(unsigned) * m.max_q;
    }
    BSHOT_CUDA_TRY(cudaMemcpyAsync(m.d_peer_gather, g.data(), sizeof(void*) * m.nranks, cudaMemcpyHostToDevice, c->stream));
    BSHOT_CUDA_TRY(cudaMemcpyAsync(m.d_peer_rq, r.data(), sizeof(void*) * m.nranks, cudaMemcpyHostToDevice, c->stream));
    BSHOT_CUDA_TRY(cudaMemcpyAsync(m.d_peer_flags, f.data(), sizeof(void*) * m.nranks, cudaMemcpyHostToDevice, c->stream));
    BSHOT_CUDA_TRY(cudaStreamSynchronize(c->stream));
    m.connected = true;
    return BSHOT_OK;
}

// the whole sharded call, asynchronous on the context stream; d_out: nq complete records (every rank gets all of them)
int hamming_match_sharded(Ctx* c, const void* d_q, size_t nq, unsigned long long global_base, bshot_cand* d_out) {
    Comm& m = c->comm;
    if (!m.connected) { set_error("sharded match: bshot_comm_create / bshot_comm_import first"); return BSHOT_E_STATE; }
    if (nq == 0) return BSHOT_OK;
    if (nq > m.max_q || nq > c->max_kp) { set_error("sharded match: %zu queries > capacity %zu", nq, std::min(m.max_q, c->max_kp)); return BSHOT_E_CAPACITY; }
    const unsigned nranks = (unsigned)m.nranks, rank = (unsigned)m.rank;
    const unsigned epoch = ++m.epoch;
    unsigned char* base = m.d_region;
    bshot_cand* gather = reinterpret_cast<bshot_cand*>(base);
    unsigned* rq = reinterpret_cast<unsigned*>(base + comm_gather_bytes(m));
    unsigned* flags = rq + m.max_q;
    unsigned* owner_q = reinterpret_cast<unsigned*>(c->d_left);
    unsigned* count = reinterpret_cast<unsigned*>(c->d_pair_count) + 1;
    unsigned* timeout_flag = reinterpret_cast<unsigned*>(c->d_pair_count) + 3;
    const unsigned qb = (unsigned)((nq + 255) / 256), mb = (unsigned)((nq * HM_MERGE_LANES + 255) / 256);
    unsigned nsplit = 1;
    // 1 + 2: shard search, merged records pushed to every rank
    if (c->n_map) {
        BSHOT_TRY(hamming_top2_partials(c, d_q, nq, c->d_map, c->n_map, global_base, nullptr, nullptr, nullptr, &nsplit));
    } else {  // empty shard: one "split" of none-candidates
        BSHOT_CUDA_TRY(cudaMemsetAsync(c->d_partial, 0xFF, sizeof(unsigned long long) * 2 * nq, c->stream));
    }
    merge_push_kernel<<<mb, 256, 0, c->stream>>>(c->d_partial, nsplit, (unsigned)nq, c->d_cand, m.d_peer_gather, m.d_peer_flags, nranks, rank, epoch,
                                                m.d_ticket, count);
    // 3: all ranks' records merged, own winners selected and gathered
    wait_merge_select_kernel<<<qb, 256, 0, c->stream>>>(gather, flags, nranks, (unsigned)nq, epoch, global_base, global_base + c->n_map,
                                                       reinterpret_cast<const uint4*>(c->d_map), d_out, reinterpret_cast<uint4*>(c->d_gather), owner_q, count,
                                                       timeout_flag);
    count_launch(c, 2);
    BSHOT_TRY(check_launch("sharded exchange kernels"));
    // 4 + 5: best query of every owned winner, pushed to every rank
    // (about nq / nranks of the rows are live: the grid of the tensor-core kernel is sized for that many)
    BSHOT_TRY(hamming_top2_partials(c, c->d_gather, nq, d_q, nq, 0, nullptr, count, nullptr, &nsplit, (nq + nranks - 1) / nranks));
    merge_push_rq_kernel<<<mb, 256, 0, c->stream>>>(c->d_partial, nsplit, (unsigned)nq, count, owner_q, m.d_peer_rq, m.d_peer_flags, nranks, rank, epoch,
                                                   m.d_ticket + 1);
    // 6
    wait_apply_rq_kernel<<<qb, 256, 0, c->stream>>>(d_out, rq, (unsigned)nq, flags, nranks, epoch, timeout_flag);
    count_launch(c, 2);
    return check_launch("sharded exchange kernels");
}

}  // namespace bshot

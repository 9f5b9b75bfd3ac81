// hamming_tc2.cu -- the tensor-core Hamming matcher as a warp-specialised pipeline (tcgen05 kind::i8, signed bytes).
//
// hamming_tc.cu showed that the distance matrix on the tensor cores is exact and ~3x the POPC kernel, but its CTA runs
// expansion -> multiply -> epilogue one after the other and spends four ALU instructions per pair on the key.  Here:
//
//   * THE MULTIPLY PRODUCES THE FINAL KEY.  A target bit becomes the byte -128, a query bit the byte 2, so a common bit adds
//     -256 = 128 * (-2); eight extra K columns carry 128 |q| (split over the query row as 127 (a + b + c) + d against target
//     bytes 127, 127, 127, 1) and 128 |t| + column (the same the other way round).  One chain of 12 K steps leaves
//         acc = 128 * (|q| + |t| - 2 <q, t>) + column = (Hamming distance << 7) | column of the tile          (<= 45183)
//     as an exact int32: the (distance, index) order of minVect (include/bshot_bits.h:6-20) inside a 128-target tile.
//   * THE EPILOGUE WORKS ON 16-BIT PAIRS.  tcgen05.ld ... pack::16b returns two columns per register; the minimum of a tile
//     costs one VIMNMX3.U16x2 per four pairs.  Only when that minimum can enter the row's running 32-bit
//     (distance << 20 | index in the chunk) top-2 is the tile scanned again for its two smallest keys per half lane
//     (three VIMNMX.U16x2 per two pairs) and merged.
//   * ROLES.  warp 0: one thread issues the MMAs; warps 1-8: expand target records into the K-major byte tile (half a row
//     per thread, next tile's records prefetched); 4 warps per query tile: epilogue.  mbarriers: full/empty per target stage,
//     full/empty per accumulator (in tensor memory, released as soon as it is in registers).  MT = 2 query tiles share
//     every expanded target tile (expansion per pair halves); MT = 1 when that pads >= 10 % fewer query rows.
//
// A query tile is 24 K chunks of 16 bytes per row (22 data, extras, zeros): 96 tensor-memory columns, or 48 KB of shared
// memory when ATM = false.  A target tile is 23 chunks (46 KB per stage): the 24th chunk of the last K step aliases the bytes
// that follow and is multiplied by the query tile's zeros.
// Output = per-split partial top-2 records like hamming_top2_kernel (hamming.cu); the merge kernels take over.
#include "common.cuh"
#include "stages.h"

namespace bshot {

constexpr int T2_KSTEPS = 12;
constexpr unsigned T2_SBO = 128;                  // bytes between 8-row groups
constexpr unsigned T2_LBO = 16 * 128;             // bytes between 16-byte K chunks: [chunk][row group 16][8 rows][16 B]
constexpr unsigned T2_A_BYTES = 24 * T2_LBO;      // 49152
constexpr unsigned T2_B_BYTES = 23 * T2_LBO;      // 47104
constexpr unsigned T2_TAIL = 2048;                // barriers + the bytes the last stage's 24th chunk aliases
constexpr unsigned T2_IDX_BITS = 20;              // target index inside a CTA's chunk
constexpr unsigned long long T2_NONE = 0xFFFFFFFFFFFFFFFFull;
constexpr int T2_XWARPS = 8;                      // expander warps (256 threads: half a target row each)

// ATM: the query tiles live in TENSOR MEMORY (96 columns each: 384 K bytes, four per column) instead of shared memory.
// Every tcgen05.mma re-reads its A operand, and the query rows never change: each epilogue thread expands its own row in
// registers and stores it once (tcgen05.st), the multiply then reads A from tensor memory ([a_tmem] operand form).  Shared
// memory only carries the target tiles (written once, read once per query tile) and holds four stages of them; two query tiles
// take turns on ONE accumulator each (the multiply of the other tile covers the time the epilogue needs to read it out).
// Measured: 2.48 -> 2.44 ms on C4 -- operand bandwidth was not the limit (the chain of multiplies alone takes 2.25 ms); kept as
// the default because it frees 96 KB of shared memory for deeper target staging.  ATM = false (BSHOT_MATCH_TC=3) keeps the
// query tiles in shared memory, two accumulators per query tile.
template <int MT, bool ATM> struct T2Cfg {
    static constexpr int STAGES = ATM ? 4 : (MT == 2 ? 2 : 3);
    static constexpr int NACC = (ATM && MT == 2) ? 1 : 2;            // accumulators per query tile
    static constexpr int THREADS = 32 * (1 + T2_XWARPS + 4 * MT);
    static constexpr unsigned SMEM = (ATM ? 0u : MT * T2_A_BYTES) + STAGES * T2_B_BYTES + T2_TAIL;
    static constexpr unsigned A_COL0 = MT * NACC * 128;              // first column of the query tiles (ATM)
    static constexpr unsigned TMEM_COLS = ATM ? 512u : MT * 2 * 128u;
};

__device__ __forceinline__ unsigned t2_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned long long t2_desc(unsigned saddr) {   // K-major, no swizzle, version 1
    return (unsigned long long)((saddr >> 4) & 0x3FFFu) | ((unsigned long long)(T2_LBO >> 4) << 16) |
           ((unsigned long long)(T2_SBO >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void t2_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}

__device__ __forceinline__ void t2_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// The K order is free as long as queries and targets agree: byte k of output word s of a 32-bit word is its bit 8 k + s,
// which is already in byte k -- a shift and a mask per four bits (the shift as IMAD.SHL on the FMA pipe, one LOP3 on the ALU
// pipe).  Target bits become 0x80 = -128, query bits 2: a common bit adds -256.
template <bool IS_A>
__device__ __forceinline__ uint4 t2_chunk(unsigned w, int h) {
    unsigned o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int s = 4 * h + k;
        if (IS_A) o[k] = (s == 0 ? (w << 1) : (w >> (s - 1))) & 0x02020202u;
        else o[k] = (w * (1u << (7 - s))) & 0x80808080u;
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// V = 127 (e0 + e1 + e2) + e3 with every e <= 127 (V <= 45183): the four bytes that meet 127, 127, 127, 1
__device__ __forceinline__ unsigned t2_extras(unsigned V) {
    const unsigned s = V / 127u, e3 = V - 127u * s;
    const unsigned e0 = min(s, 127u), e1 = min(s - e0, 127u), e2 = s - e0 - e1;
    return e0 | (e1 << 8) | (e2 << 16) | (e3 << 24);
}

template <bool IS_A>
__device__ __forceinline__ void t2_expand_words(unsigned char* rowbase, const unsigned* w, int first, int count) {
#pragma unroll
    for (int i = 0; i < count; ++i) {
        *reinterpret_cast<uint4*>(rowbase + (unsigned)(2 * (first + i)) * T2_LBO) = t2_chunk<IS_A>(w[i], 0);
        *reinterpret_cast<uint4*>(rowbase + (unsigned)(2 * (first + i) + 1) * T2_LBO) = t2_chunk<IS_A>(w[i], 1);
    }
}

__device__ __forceinline__ void t2_top2x2(unsigned& k1, unsigned& k2, unsigned key) {
    const unsigned hi = __vmaxu2(k1, key);
    k1 = __vminu2(k1, key);
    k2 = __vminu2(k2, hi);
}

__device__ __forceinline__ void t2_insert(unsigned& k1, unsigned& k2, unsigned key) {
    const unsigned hi = max(k1, key);
    k1 = min(k1, key);
    k2 = min(k2, hi);
}

#define T2_LD32_PACK(r, o, taddr)                                                                                            \
    asm volatile(                                                                                                           \
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "                                                                  \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                            \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                            \
        : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]),     \
          "=r"(r[o + 7]), "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]), "=r"(r[o + 12]), "=r"(r[o + 13]), \
          "=r"(r[o + 14]), "=r"(r[o + 15]), "=r"(r[o + 16]), "=r"(r[o + 17]), "=r"(r[o + 18]), "=r"(r[o + 19]),              \
          "=r"(r[o + 20]), "=r"(r[o + 21]), "=r"(r[o + 22]), "=r"(r[o + 23]), "=r"(r[o + 24]), "=r"(r[o + 25]),              \
          "=r"(r[o + 26]), "=r"(r[o + 27]), "=r"(r[o + 28]), "=r"(r[o + 29]), "=r"(r[o + 30]), "=r"(r[o + 31])               \
        : "r"(taddr) : "memory")

template <int MT, bool ATM>
__global__ void __launch_bounds__(T2Cfg<MT, ATM>::THREADS, 1)
hamming_tc2_kernel(const uint4* __restrict__ q, unsigned nq, const unsigned* __restrict__ nq_dev, const uint4* __restrict__ t, unsigned nt,
                   const unsigned* __restrict__ nt_dev, unsigned chunk, unsigned long long global_base, unsigned long long* __restrict__ partial) {
    using Cfg = T2Cfg<MT, ATM>;
    constexpr int STAGES = Cfg::STAGES, NACC = Cfg::NACC;
    extern __shared__ __align__(128) unsigned char t2_smem_raw[];
    unsigned char* sA = t2_smem_raw;
    unsigned char* sB = t2_smem_raw + (ATM ? 0u : MT * T2_A_BYTES);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sB + STAGES * T2_B_BYTES);
    // full[STAGES] | empty[STAGES] | accfull[MT][NACC] | accempty[MT][NACC]
    const unsigned bar_full = t2_smem(bars), bar_empty = bar_full + 8u * STAGES, bar_accfull = bar_empty + 8u * STAGES,
                   bar_accempty = bar_accfull + 8u * NACC * MT;
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 2 * STAGES + 2 * NACC * MT);
    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const unsigned nq_live = nq_dev ? min(nq, *nq_dev) : nq, nt_live = nt_dev ? min(nt, *nt_dev) : nt;
    const unsigned q0 = blockIdx.x * (128u * MT), t0 = blockIdx.y * chunk;
    const unsigned t1 = min(nt_live, t0 + chunk);
    const unsigned ntiles = t1 > t0 ? (t1 - t0 + 127u) / 128u : 0u;

    if (q0 >= nq_live) {   // a query block beyond the device-side count (reverse pass of the sharded call: few owned winners)
        for (unsigned x = tid; x < 128u * MT; x += Cfg::THREADS) {
            if (q0 + x < nq) {
                unsigned long long* p = partial + ((size_t)blockIdx.y * nq + q0 + x) * 2;
                p[0] = T2_NONE;
                p[1] = T2_NONE;
            }
        }
        return;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(t2_smem(tmem_slot)), "r"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_full + 8u * s), "r"((unsigned)T2_XWARPS) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_empty + 8u * s), "r"(1u) : "memory");
        }
        for (int a = 0; a < NACC * MT; ++a) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_accfull + 8u * a), "r"(1u) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_accempty + 8u * a), "r"(4u) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the query tiles, expanded once: expander thread x owns row x of the CTA's 128 MT queries
    if (!ATM && warp >= 1 && warp <= T2_XWARPS) {
        for (unsigned x = tid - 32u; x < 128u * MT; x += 32u * T2_XWARPS) {
            const unsigned qi = q0 + x, row = x & 127u;
            unsigned char* rowbase = sA + (x >> 7) * T2_A_BYTES + (row >> 3) * T2_SBO + (row & 7u) * 16u;
            uint4 a = make_uint4(0, 0, 0, 0), b = a, c = a;
            if (qi < nq_live) { a = __ldg(q + 3 * (size_t)qi); b = __ldg(q + 3 * (size_t)qi + 1); c = __ldg(q + 3 * (size_t)qi + 2); }
            const unsigned w[11] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z};
            unsigned pc = 0;
#pragma unroll
            for (int i = 0; i < 11; ++i) pc += __popc(w[i]);
            t2_expand_words<true>(rowbase, w, 0, 11);
            *reinterpret_cast<uint4*>(rowbase + 22u * T2_LBO) = make_uint4(0x017F7F7Fu, t2_extras(128u * pc), 0u, 0u);
            *reinterpret_cast<uint4*>(rowbase + 23u * T2_LBO) = make_uint4(0u, 0u, 0u, 0u);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (ATM) {   // the tensor-memory address must be known before the query rows can be stored
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (warp > T2_XWARPS) {   // epilogue thread = query row = tensor-memory lane: expand the row in registers, eight columns a store
            const unsigned e = warp - (1u + T2_XWARPS), mt = e >> 2, quarter = warp & 3u;
            const unsigned qi = q0 + mt * 128u + quarter * 32u + lane;
            uint4 a = make_uint4(0, 0, 0, 0), b = a, c = a;
            if (qi < nq_live) { a = __ldg(q + 3 * (size_t)qi); b = __ldg(q + 3 * (size_t)qi + 1); c = __ldg(q + 3 * (size_t)qi + 2); }
            const unsigned w[11] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z};
            unsigned pc = 0;
            const unsigned acol = *tmem_slot + ((quarter * 32u) << 16) + Cfg::A_COL0 + mt * 96u;
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                uint4 lo, hi;
                if (i < 11) {
                    pc += __popc(w[i]);
                    lo = t2_chunk<true>(w[i], 0);
                    hi = t2_chunk<true>(w[i], 1);
                } else {
                    lo = make_uint4(0x017F7F7Fu, t2_extras(128u * pc), 0u, 0u);
                    hi = make_uint4(0u, 0u, 0u, 0u);
                }
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                             ::"r"(acol + 8u * i), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = *tmem_slot;

    if (warp == 0) {
        // ---- MMA issue ------------------------------------------------------------------------------------------------
        if (lane == 0) {
            // kind::i8: D = S32, A = B = signed 8 bit, both K-major, N = 128, M = 128
            const unsigned idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
            for (unsigned i = 0; i < ntiles; ++i) {
                const unsigned s = i % STAGES, ph = (i / STAGES) & 1u, b = i % NACC, aph = (i / NACC) & 1u;
                t2_wait(bar_full + 8u * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned long long descB = t2_desc(t2_smem(sB + s * T2_B_BYTES));
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    t2_wait(bar_accempty + 8u * (NACC * mt + b), aph ^ 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned long long descA = t2_desc(t2_smem(sA + mt * T2_A_BYTES));
                    const unsigned a_tmem = tmem + Cfg::A_COL0 + (unsigned)mt * 96u;
                    const unsigned d_tmem = tmem + (unsigned)(NACC * mt + b) * 128u;
#pragma unroll
                    for (int ks = 0; ks < T2_KSTEPS; ++ks) {
                        const unsigned long long da = descA + (unsigned long long)((2u * T2_LBO * ks) >> 4);
                        const unsigned long long db = descB + (unsigned long long)((2u * T2_LBO * ks) >> 4);
                        const unsigned acc = ks > 0 ? 1u : 0u;
                        if (ATM)
                            asm volatile(
                                "{\n\t.reg .pred p;\n\t"
                                "setp.ne.b32 p, %4, 0;\n\t"
                                "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                                ::"r"(d_tmem), "r"(a_tmem + 8u * ks), "l"(db), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
                        else
                        asm volatile(
                            "{\n\t.reg .pred p;\n\t"
                            "setp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                            ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_accfull + 8u * (NACC * mt + b)) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_empty + 8u * s) : "memory");
            }
        }
        __syncwarp();
    } else if (warp <= T2_XWARPS) {
        // ---- expanders: half a target row per thread ----------------------------------------------------------------------
        const unsigned x = tid - 32u, row = x & 127u, half = x >> 7;
        const unsigned rowoff = (row >> 3) * T2_SBO + (row & 7u) * 16u;
        uint4 ra = make_uint4(0, 0, 0, 0), rb = ra, rc = ra;
        auto fetch = [&](unsigned ti) {
            ra = make_uint4(0, 0, 0, 0); rb = ra; rc = ra;
            if (ti < t1) {
                const uint4* p = t + 3 * (size_t)ti;
                if (half == 0) { ra = __ldg(p); rb = __ldg(p + 1); } else { ra = __ldg(p); rb = __ldg(p + 1); rc = __ldg(p + 2); }
            }
        };
        fetch(t0 + row);
        for (unsigned i = 0; i < ntiles; ++i) {
            const unsigned s = i % STAGES, ph = (i / STAGES) & 1u;
            const uint4 a = ra, b = rb, c = rc;
            fetch(t0 + (i + 1u) * 128u + row);   // the next tile's records travel while this one is expanded
            t2_wait(bar_empty + 8u * s, ph ^ 1u);
            unsigned char* rowbase = sB + s * T2_B_BYTES + rowoff;
#ifdef T2_DBG_NOEXPAND   // timing experiment: the multiply runs on whatever the stage holds
            if (a.x == 0x12345u) *reinterpret_cast<uint4*>(rowbase) = a;
#else
            if (half == 0) {
                const unsigned w[6] = {a.x, a.y, a.z, a.w, b.x, b.y};
                t2_expand_words<false>(rowbase, w, 0, 6);
            } else {
                const unsigned w[5] = {b.z, b.w, c.x, c.y, c.z};
                t2_expand_words<false>(rowbase, w, 6, 5);
                const unsigned pc = ((__popc(a.x) + __popc(a.y)) + (__popc(a.z) + __popc(a.w))) + ((__popc(b.x) + __popc(b.y)) + (__popc(b.z) + __popc(b.w))) +
                                    ((__popc(c.x) + __popc(c.y)) + __popc(c.z));
                // rows beyond the range are all zeros and still carry their column: the epilogue masks them by it
                *reinterpret_cast<uint4*>(rowbase + 22u * T2_LBO) = make_uint4(t2_extras(128u * pc + row), 0x017F7F7Fu, 0u, 0u);
            }
#endif
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor-core reads
            __syncwarp();
            if (lane == 0) t2_arrive(bar_full + 8u * s);
        }
    } else {
        // ---- epilogue: one thread per query row ------------------------------------------------------------------------------
        const unsigned e = warp - (1u + T2_XWARPS), mt = e >> 2, quarter = warp & 3u;   // a warp reads the TMEM lanes 32 (warp % 4) ..
        const unsigned rowq = mt * 128u + quarter * 32u + lane;
        unsigned K1 = 0xFFFFFFFFu, K2 = 0xFFFFFFFFu;
        for (unsigned i = 0; i < ntiles; ++i) {
            const unsigned b = i % NACC, aph = (i / NACC) & 1u;
            t2_wait(bar_accfull + 8u * (NACC * mt + b), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            unsigned r[64];
            const unsigned taddr = tmem + ((quarter * 32u) << 16) + (NACC * mt + b) * 128u;
#ifdef T2_DBG_NOEPI      // timing experiment: the accumulator is released unread
#pragma unroll
            for (int j = 0; j < 64; ++j) r[j] = taddr + j;
#else
            T2_LD32_PACK(r, 0, taddr);
            T2_LD32_PACK(r, 32, taddr + 64u);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#endif
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) t2_arrive(bar_accempty + 8u * (NACC * mt + b));   // the accumulator is in registers: the next multiply may start
            const unsigned nv = t1 - (t0 + i * 128u);
            if (nv < 128u) {   // last tile of the target range: columns beyond it never win
#pragma unroll
                for (int j = 0; j < 64; ++j) r[j] |= ((r[j] & 127u) >= nv ? 0xFFFFu : 0u) | (((r[j] >> 16) & 127u) >= nv ? 0xFFFF0000u : 0u);
            }
            // most tiles cannot change a row's top 2: the minimum alone (one VIMNMX3 per two registers) decides
            unsigned m0 = r[0], m1 = r[1];
#pragma unroll
            for (int j = 2; j < 64; j += 2) {
                m0 = __vminu2(m0, r[j]);
                m1 = __vminu2(m1, r[j + 1]);
            }
            m0 = __vminu2(m0, m1);
            const unsigned best = min(m0 & 0xFFFFu, m0 >> 16);
            if ((best >> 7) < (K2 >> T2_IDX_BITS)) {   // equal distance, later index: cannot enter
                unsigned a1 = 0xFFFFFFFFu, a2 = 0xFFFFFFFFu, b1 = 0xFFFFFFFFu, b2 = 0xFFFFFFFFu;
#pragma unroll
                for (int j = 0; j < 64; j += 2) {
                    t2_top2x2(a1, a2, r[j]);
                    t2_top2x2(b1, b2, r[j + 1]);
                }
                const unsigned hi = __vmaxu2(a1, b1);
                const unsigned k1 = __vminu2(a1, b1), k2 = __vminu2(__vminu2(a2, b2), hi);
                const unsigned c[4] = {k1 & 0xFFFFu, k1 >> 16, k2 & 0xFFFFu, k2 >> 16};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned key = c[u] == 0xFFFFu ? 0xFFFFFFFFu : (((c[u] >> 7) << T2_IDX_BITS) | (i * 128u + (c[u] & 127u)));
                    t2_insert(K1, K2, key);
                }
            }
        }
        const unsigned qi = q0 + rowq;
        if (qi < nq) {
            const unsigned long long gb = global_base + t0;
            unsigned long long o1 = T2_NONE, o2 = T2_NONE;
            if (K1 != 0xFFFFFFFFu) o1 = ((unsigned long long)(K1 >> T2_IDX_BITS) << 32) | (gb + (K1 & ((1u << T2_IDX_BITS) - 1u)));
            if (K2 != 0xFFFFFFFFu) o2 = ((unsigned long long)(K2 >> T2_IDX_BITS) << 32) | (gb + (K2 & ((1u << T2_IDX_BITS) - 1u)));
            unsigned long long* p = partial + ((size_t)blockIdx.y * nq + qi) * 2;
            p[0] = o1;
            p[1] = o2;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS) : "memory");
}

template <int MT, bool ATM>
static int t2_launch(Ctx* c, const void* d_q, size_t nq, const void* d_t, size_t nt, unsigned long long global_base, const unsigned* d_nq,
                     const unsigned* d_nt, unsigned* nsplit_out, size_t live_q) {
    using Cfg = T2Cfg<MT, ATM>;
    static bool attr_set = false;
    if (!attr_set) {
        BSHOT_CUDA_TRY(cudaFuncSetAttribute(hamming_tc2_kernel<MT, ATM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_set = true;
    }
    const size_t qblocks = (nq + 128 * MT - 1) / (128 * MT);
    const size_t ntile = (nt + 127) / 128;
    const size_t cap_splits = std::min<size_t>(c->partial_cap / (nq * 2), 65535);
    if (cap_splits == 0) { set_error("hamming_tc2: partial buffer too small for %zu queries", nq); return BSHOT_E_CAPACITY; }
    // one CTA per SM: the grid fills k whole waves as closely as it can; a CTA pays about three tiles' worth of prologue
    // live_q: how many of the nq rows the caller expects to be live under the device-side count (dead blocks leave at once)
    const size_t qblocks_live = std::max<size_t>(1, std::min(qblocks, (std::min(live_q, nq) + 128 * MT - 1) / (128 * MT)));
    const size_t sms = (size_t)c->sm_count;
    size_t best_split = 1;
    double best_cost = 1e30;
    for (size_t k = 1; k <= 16; ++k) {
        size_t ns = std::max<size_t>(1, k * sms / qblocks_live);
        ns = std::min(ns, std::min(cap_splits, std::max<size_t>(1, ntile)));
        const size_t tiles = (ntile + ns - 1) / ns;                       // tiles per CTA
        const size_t waves = (qblocks_live * ((ntile + tiles - 1) / tiles) + sms - 1) / sms;
        const double cost = (double)waves * ((double)tiles + 3.0);          // ~3 tiles' worth of prologue per CTA
        if (cost < best_cost) { best_cost = cost; best_split = ns; }
    }
    size_t chunk = ((ntile + best_split - 1) / best_split) * 128;
    if (chunk > (1u << T2_IDX_BITS)) chunk = 1u << T2_IDX_BITS;
    const size_t nsplit = std::max<size_t>(1, (nt + chunk - 1) / chunk);
    if (nsplit > cap_splits) { set_error("hamming_tc2: %zu splits exceed the partial buffer", nsplit); return BSHOT_E_CAPACITY; }
    const dim3 grid((unsigned)qblocks, (unsigned)nsplit);
    hamming_tc2_kernel<MT, ATM><<<grid, Cfg::THREADS, Cfg::SMEM, c->stream>>>(reinterpret_cast<const uint4*>(d_q), (unsigned)nq, d_nq, reinterpret_cast<const uint4*>(d_t),
                                                                      (unsigned)nt, d_nt, (unsigned)chunk, global_base, c->d_partial);
    count_launch(c);
    *nsplit_out = (unsigned)nsplit;
    return check_launch("hamming_tc2_kernel");
}

// see hamming_preload_sharded (hamming.cu): every kernel a sharded call may launch is loaded before the first call
int hamming_tc2_preload() {
    cudaFuncAttributes a;
    BSHOT_CUDA_TRY(cudaFuncGetAttributes(&a, hamming_tc2_kernel<1, true>));
    BSHOT_CUDA_TRY(cudaFuncGetAttributes(&a, hamming_tc2_kernel<2, true>));
    BSHOT_CUDA_TRY(cudaFuncGetAttributes(&a, hamming_tc2_kernel<1, false>));
    BSHOT_CUDA_TRY(cudaFuncGetAttributes(&a, hamming_tc2_kernel<2, false>));
    return BSHOT_OK;
}

// per-split top-2 candidates in c->d_partial ([nsplit][nq][2]) like hamming_top2_partials (hamming.cu)
int hamming_tc2_partials(Ctx* c, const void* d_q, size_t nq, const void* d_t, size_t nt, unsigned long long global_base, const unsigned* d_nq,
                         const unsigned* d_nt, unsigned* nsplit_out, size_t live_q) {
    if (live_q == 0 || !d_nq) live_q = nq;
    if (nq > 0xFFFFFFFFull || nt > 0xFFFFFFFFull || global_base + nt > 0x100000000ull) { set_error("hamming_tc2: sizes exceed 32-bit index range"); return BSHOT_E_INVALID; }
    if (c->match_tc == 3) {   // query tiles in shared memory (kept for comparison; never chosen by size)
        if (nq > 128) return t2_launch<2, false>(c, d_q, nq, d_t, nt, global_base, d_nq, d_nt, nsplit_out, live_q);
        return t2_launch<1, false>(c, d_q, nq, d_t, nt, global_base, d_nq, d_nt, nsplit_out, live_q);
    }
    // two query tiles per CTA halve the expansion work per pair; one tile per CTA when that pads >= 10 % fewer rows
    const size_t rows1 = (nq + 127) / 128 * 128, rows2 = (nq + 255) / 256 * 256;
    if (rows1 * 10 > rows2 * 9) return t2_launch<2, true>(c, d_q, nq, d_t, nt, global_base, d_nq, d_nt, nsplit_out, live_q);
    return t2_launch<1, true>(c, d_q, nq, d_t, nt, global_base, d_nq, d_nt, nsplit_out, live_q);
}

}  // namespace bshot

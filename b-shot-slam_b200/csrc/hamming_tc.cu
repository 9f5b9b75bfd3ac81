// hamming_tc.cu -- the Hamming distance matrix on the 5th-generation tensor cores (tcgen05, kind::i8).
//
// popcount(q xor t) = |q| + |t| - 2 <q, t> for 0/1 vectors: the 352-bit descriptors are expanded to 352 bytes of 0/1 in
// shared memory (K-major, no swizzle: 8 x 16-byte core matrices) and one tcgen05.mma chain (11 steps of K = 32) leaves the
// 128 x 128 dot products of a query tile and a target tile in tensor memory as exact int32.  The epilogue (one thread per
// query row, tcgen05.ld 32x32b) turns every accumulator into a packed key ((|t| + 512 - 2 dot) << 20 | local target index)
// with one IMAD and keeps the two smallest with three min/max -- the same (distance, index) order, hence the same
// first-minimum winners, as minVect (include/bshot_bits.h:6-20) over src/lidar_odometry.cpp:217-232.  |q| is the same for a
// whole row and joins at the end.  Output = the per-split partial top-2 records of hamming_top2_kernel (hamming.cu), so the
// existing merge kernels take over.
//
// One CTA = 128 queries x one chunk of targets, 128 threads; two CTAs per SM overlap each other's expansion / MMA / epilogue.
#include "common.cuh"
#include "stages.h"

#ifndef BSHOT_TC_LUT
#define BSHOT_TC_LUT 0     // 1: bits -> bytes through a 256-entry shared-memory table instead of multiplies
#endif
#ifndef BSHOT_TC_GROUP
#define BSHOT_TC_GROUP 0   // 1: eight keys share one "can any of them enter the top 2" test (measured slower: 5.7 vs 4.96 ms on C4)
#endif

namespace bshot {

constexpr int TC_M = 128;                       // queries per CTA (UMMA M)
constexpr int TC_N = 128;                       // targets per tile (UMMA N)
constexpr int TC_KSTEPS = 11;                   // 352 bytes / 32 per MMA
constexpr unsigned TC_SBO = 128;                // bytes between 8-row groups
constexpr unsigned TC_LBO = 16 * 128;           // bytes between 16-byte K chunks: [kc 22][row group 16][8 rows][16 B]
constexpr unsigned TC_TILE_BYTES = 22 * TC_LBO; // 45056
constexpr unsigned TC_SMEM = 2 * TC_TILE_BYTES + TC_N * 4 + 64 + 256 * 8;
constexpr unsigned TC_IDX_BITS = 20;            // local target index inside a CTA's chunk
constexpr unsigned TC_BIAS = 512;               // keeps |t| - 2 dot non-negative
constexpr unsigned long long TC_NONE = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (start address, leading = K direction, stride = row groups, version 1)
__device__ __forceinline__ unsigned long long tc_desc(unsigned saddr) {
    return (unsigned long long)((saddr >> 4) & 0x3FFFu) | ((unsigned long long)(TC_LBO >> 4) << 16) |
           ((unsigned long long)(TC_SBO >> 4) << 32) | (1ull << 46);
}

// 352 bits (three uint4: words 0..10 carry bits) -> 352 bytes of 0/1, row `row` of a tile; returns the popcount.
// lut[b] = the eight 0/1 bytes of byte b (256 x 8 B in shared memory)
__device__ __forceinline__ unsigned tc_expand_row(unsigned char* tile, unsigned row, const uint4 a, const uint4 b, const uint4 c, const uint2* lut) {
    const unsigned w[11] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z};
    unsigned char* base = tile + (row >> 3) * TC_SBO + (row & 7u) * 16u;
    unsigned pc = 0;
#pragma unroll
    for (int i = 0; i < 11; ++i) {
        pc += __popc(w[i]);
#if BSHOT_TC_LUT
        const uint2 e0 = lut[w[i] & 0xFFu], e1 = lut[(w[i] >> 8) & 0xFFu], e2 = lut[(w[i] >> 16) & 0xFFu], e3 = lut[w[i] >> 24];
        *reinterpret_cast<uint4*>(base + (unsigned)(2 * i) * TC_LBO) = make_uint4(e0.x, e0.y, e1.x, e1.y);       // bits 0..15: one 16-byte K chunk
        *reinterpret_cast<uint4*>(base + (unsigned)(2 * i + 1) * TC_LBO) = make_uint4(e2.x, e2.y, e3.x, e3.y);   // bits 16..31
#else
#pragma unroll
        for (int h = 0; h < 2; ++h) {   // 16 bits -> one 16-byte K chunk; a nibble n becomes four 0/1 bytes: (n * 0x00204081) & 0x01010101
            const unsigned v = (w[i] >> (16 * h)) & 0xFFFFu;
            uint4 o;
            o.x = ((v & 0xFu) * 0x00204081u) & 0x01010101u;
            o.y = (((v >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
            o.z = (((v >> 8) & 0xFu) * 0x00204081u) & 0x01010101u;
            o.w = (((v >> 12) & 0xFu) * 0x00204081u) & 0x01010101u;
            *reinterpret_cast<uint4*>(base + (unsigned)(2 * i + h) * TC_LBO) = o;
        }
#endif
    }
    return pc;
}

__device__ __forceinline__ void tc_mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}

__global__ void __launch_bounds__(TC_M, 2)
hamming_tc_kernel(const uint4* __restrict__ q, unsigned nq, const unsigned* __restrict__ nq_dev, const uint4* __restrict__ t, unsigned nt,
                  const unsigned* __restrict__ nt_dev, unsigned chunk, unsigned long long global_base, unsigned long long* __restrict__ partial) {
    extern __shared__ __align__(128) unsigned char tc_smem[];
    unsigned char* sA = tc_smem;
    unsigned char* sB = tc_smem + TC_TILE_BYTES;
    unsigned* tbase = reinterpret_cast<unsigned*>(tc_smem + 2 * TC_TILE_BYTES);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(tc_smem + 2 * TC_TILE_BYTES + TC_N * 4);
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bar + 1);
    uint2* lut = reinterpret_cast<uint2*>(tc_smem + 2 * TC_TILE_BYTES + TC_N * 4 + 64);
    const unsigned tid = threadIdx.x, warp = tid >> 5;
    const unsigned nq_live = nq_dev ? min(nq, *nq_dev) : nq, nt_live = nt_dev ? min(nt, *nt_dev) : nt;
    const unsigned q0 = blockIdx.x * TC_M, t0 = blockIdx.y * chunk;
    const unsigned t1 = min(nt_live, t0 + chunk);
    const unsigned bar_addr = smem_u32(bar);

    if (warp == 0) {  // tensor-memory columns for one 128 x 128 int32 accumulator
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_addr), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (unsigned v = tid; v < 256; v += TC_M)
        lut[v] = make_uint2(((v & 0xFu) * 0x00204081u) & 0x01010101u, (((v >> 4) & 0xFu) * 0x00204081u) & 0x01010101u);
    __syncthreads();
    // the query tile, expanded once
    unsigned pq = 0;
    {
        const unsigned qi = q0 + tid;
        uint4 a = make_uint4(0, 0, 0, 0), b = a, c = a;
        if (qi < nq_live) { a = __ldg(q + 3 * (size_t)qi); b = __ldg(q + 3 * (size_t)qi + 1); c = __ldg(q + 3 * (size_t)qi + 2); }
        pq = tc_expand_row(sA, tid, a, b, c, lut);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = *tmem_slot;
    const unsigned long long descA = tc_desc(smem_u32(sA)), descB = tc_desc(smem_u32(sB));
    // kind::i8: D = S32, A = B = unsigned 8 bit, both K-major, N = 128, M = 128
    const unsigned idesc = (2u << 4) | ((unsigned)(TC_N >> 3) << 17) | ((unsigned)(TC_M >> 4) << 24);
    const unsigned neg2 = 0u - (1u << (TC_IDX_BITS + 1));   // key = tbase - 2 * dot << 20

    unsigned k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu, parity = 0;
    // records of the tile after the current one travel while this one is multiplied and scanned
    uint4 ra = make_uint4(0, 0, 0, 0), rb = ra, rc = ra;
    if (t0 + tid < t1) { ra = __ldg(t + 3 * (size_t)(t0 + tid)); rb = __ldg(t + 3 * (size_t)(t0 + tid) + 1); rc = __ldg(t + 3 * (size_t)(t0 + tid) + 2); }
    for (unsigned tile = t0; tile < t1; tile += TC_N) {
        {   // the target tile (rows beyond the range: zero bytes, a key that never wins)
            const unsigned ti = tile + tid;
            const bool valid = ti < t1;
            const unsigned pt = tc_expand_row(sB, tid, ra, rb, rc, lut);
            tbase[tid] = valid ? (((pt + TC_BIAS) << TC_IDX_BITS) | (ti - t0)) : 0xFFFFFFFFu;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor-core reads
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < TC_KSTEPS; ++ks) {
                const unsigned long long da = descA + (unsigned long long)((2u * TC_LBO * ks) >> 4);
                const unsigned long long db = descB + (unsigned long long)((2u * TC_LBO * ks) >> 4);
                const unsigned acc = ks > 0 ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                    ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
        }
        {   // next tile's records
            const unsigned ti = tile + TC_N + tid;
            ra = make_uint4(0, 0, 0, 0); rb = ra; rc = ra;
            if (ti < t1) { ra = __ldg(t + 3 * (size_t)ti); rb = __ldg(t + 3 * (size_t)ti + 1); rc = __ldg(t + 3 * (size_t)ti + 2); }
        }
        tc_mbar_wait(bar_addr, parity);
        parity ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // row `tid` of the accumulator lives in TMEM lane tid: warp w reads lanes 32 w .. 32 w + 31
#pragma unroll
        for (int c0 = 0; c0 < TC_N; c0 += 32) {
            unsigned d[32];
            const unsigned taddr = tmem + ((warp * 32u) << 16) + (unsigned)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]),
                  "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]), "=r"(d[16]), "=r"(d[17]), "=r"(d[18]),
                  "=r"(d[19]), "=r"(d[20]), "=r"(d[21]), "=r"(d[22]), "=r"(d[23]), "=r"(d[24]), "=r"(d[25]), "=r"(d[26]), "=r"(d[27]),
                  "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])
                : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const uint4* tb4 = reinterpret_cast<const uint4*>(tbase + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {   // eight keys at a time: their minimum decides whether any of them can enter the top 2
                const uint4 ta = tb4[2 * j], tb = tb4[2 * j + 1];
                unsigned key[8];
                key[0] = d[8 * j] * neg2 + ta.x; key[1] = d[8 * j + 1] * neg2 + ta.y; key[2] = d[8 * j + 2] * neg2 + ta.z; key[3] = d[8 * j + 3] * neg2 + ta.w;
                key[4] = d[8 * j + 4] * neg2 + tb.x; key[5] = d[8 * j + 5] * neg2 + tb.y; key[6] = d[8 * j + 6] * neg2 + tb.z; key[7] = d[8 * j + 7] * neg2 + tb.w;
#if BSHOT_TC_GROUP
                const unsigned m = min(min(min(key[0], key[1]), min(key[2], key[3])), min(min(key[4], key[5]), min(key[6], key[7])));
                if (m < k2)
#endif
                {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const unsigned hi = max(k1, key[e]);
                        k1 = min(k1, key[e]);
                        k2 = min(k2, hi);
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();   // every row is done with the accumulator, tbase and the target tile
    }
    {
        const unsigned qi = q0 + tid;
        if (qi < nq) {
            const unsigned long long gb = global_base + t0;
            unsigned long long o1 = TC_NONE, o2 = TC_NONE;
            if (k1 != 0xFFFFFFFFu) o1 = ((unsigned long long)((k1 >> TC_IDX_BITS) - TC_BIAS + pq) << 32) | (gb + (k1 & ((1u << TC_IDX_BITS) - 1u)));
            if (k2 != 0xFFFFFFFFu) o2 = ((unsigned long long)((k2 >> TC_IDX_BITS) - TC_BIAS + pq) << 32) | (gb + (k2 & ((1u << TC_IDX_BITS) - 1u)));
            unsigned long long* p = partial + ((size_t)blockIdx.y * nq + qi) * 2;
            p[0] = o1;
            p[1] = o2;
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

// per-split top-2 candidates in c->d_partial ([nsplit][nq][2]) like hamming_top2_partials (hamming.cu)
int hamming_tc_partials(Ctx* c, const void* d_q, size_t nq, const void* d_t, size_t nt, unsigned long long global_base, const unsigned* d_nq,
                        const unsigned* d_nt, unsigned* nsplit_out) {
    if (nq > 0xFFFFFFFFull || nt > 0xFFFFFFFFull || global_base + nt > 0x100000000ull) { set_error("hamming_tc: sizes exceed 32-bit index range"); return BSHOT_E_INVALID; }
    static bool attr_set = false;
    if (!attr_set) {
        BSHOT_CUDA_TRY(cudaFuncSetAttribute(hamming_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        attr_set = true;
    }
    const size_t qblocks = (nq + TC_M - 1) / TC_M;
    const size_t slots = (size_t)c->sm_count * 2;
    size_t want = std::max<size_t>(1, (2 * slots + qblocks - 1) / qblocks);   // about two waves
    const size_t cap_splits = c->partial_cap / (nq * 2);
    if (cap_splits == 0) { set_error("hamming_tc: partial buffer too small for %zu queries", nq); return BSHOT_E_CAPACITY; }
    want = std::min(want, std::min<size_t>(cap_splits, 65535));
    size_t chunk = (nt + want - 1) / want;
    chunk = std::max<size_t>((chunk + TC_N - 1) / TC_N * TC_N, 4 * TC_N);
    if (chunk > (1u << TC_IDX_BITS)) chunk = 1u << TC_IDX_BITS;
    const size_t nsplit = std::max<size_t>(1, (nt + chunk - 1) / chunk);
    if (nsplit > cap_splits || nsplit > 65535) { set_error("hamming_tc: %zu splits exceed the partial buffer", nsplit); return BSHOT_E_CAPACITY; }
    const dim3 grid((unsigned)qblocks, (unsigned)nsplit);
    hamming_tc_kernel<<<grid, TC_M, TC_SMEM, c->stream>>>(reinterpret_cast<const uint4*>(d_q), (unsigned)nq, d_nq, reinterpret_cast<const uint4*>(d_t), (unsigned)nt,
                                                       d_nt, (unsigned)chunk, global_base, c->d_partial);
    count_launch(c);
    *nsplit_out = (unsigned)nsplit;
    return check_launch("hamming_tc_kernel");
}

}  // namespace bshot

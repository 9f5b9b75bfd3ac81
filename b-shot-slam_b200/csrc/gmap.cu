// gmap.cu -- GPU-resident global keypoint map (SURVEY 8a row a9 / 8f "next" #2).
//
// Replaces the reference's host map on the hot path: Keypoint::createKeypoint (src/keypoint.cpp:23-32: position snapped
// to the 10 mm lattice by truncation), Map::addKeypoint (src/mymap.cpp:4-26: 10 m blocks; a keypoint is rejected when a
// stored keypoint of its block lies within 800 mm and is at least as salient; otherwise inserted, or overwrites the entry
// at exactly the same position), LidarOdometry::updateMap (src/lidar_odometry.cpp:343-358: every keypoint of the frame,
// in keypoint order, transformed by the frame pose) and Map::getKeypoints (src/mymap.cpp:28-74: all keypoints of the
// blocks in the +-range cube around a position, blocks visited x-outermost / z-innermost).  The gathered descriptors
// land straight in the matcher's target buffer, followed by the reference frame's descriptors
// (src/lidar_odometry.cpp:197-206) -- no host re-assembly, no PCIe upload of the target set.
//
// Order inside a block: the reference iterates an unordered_map (implementation defined); here, as in the
// host mirror (host/mymap.h) and the oracle (orc_map_*), it is INSERTION order (an overwrite keeps its slot).
//
// Layout: entries are append-only SoA arrays (position + seg-ratio as float4, 48-byte descriptor); a block is a hash
// table slot {key, count, first chunk, last chunk} plus a chain of 32-index chunks, so a warp reads a block 32 entries
// at a time.  Admission is sequential per block (the rule is order dependent) and parallel across blocks: one warp per
// block touched by the frame walks the frame's keypoints in order.
#include "common.cuh"
#include "stages.h"

namespace bshot {

constexpr unsigned GM_PREC = 10000;          // block edge (mm), src/mymap.h:49
constexpr float GM_LATTICE = 10.0f;          // position lattice (mm), src/keypoint.cpp:25
constexpr float GM_REJECT_MM = 800.0f;       // src/mymap.cpp:17
constexpr unsigned long long GM_OCC = 1ull << 63;

struct GmapBlock {
    unsigned long long key;   // GM_OCC | block id (21 bits per axis), 0 = free
    unsigned count, first, last, touched;
};

// block id of a position: every axis rounded to the 10 m lattice, 21 bits each (src/mymap.cpp:103-112)
__host__ __device__ __forceinline__ unsigned long long gm_block_id(float x, float y, float z) {
    const int gx = (int)roundf(x / (float)GM_PREC) * (int)GM_PREC, gy = (int)roundf(y / (float)GM_PREC) * (int)GM_PREC,
              gz = (int)roundf(z / (float)GM_PREC) * (int)GM_PREC;
    return (((unsigned long long)(long long)gx & 0x1FFFFFull) << 42) | (((unsigned long long)(long long)gy & 0x1FFFFFull) << 21) |
           ((unsigned long long)(long long)gz & 0x1FFFFFull);
}

__device__ __forceinline__ unsigned gm_hash(unsigned long long k, unsigned mask) {
    k ^= k >> 33; k *= 0xFF51AFD7ED558CCDull; k ^= k >> 33;
    return (unsigned)k & mask;
}

// slot of the block, optionally creating it; 0xFFFFFFFF = absent (lookup) / table full (insert)
__device__ __forceinline__ unsigned gm_find(GmapBlock* tab, unsigned mask, unsigned long long id, bool create) {
    const unsigned long long key = GM_OCC | id;
    unsigned h = gm_hash(id, mask);
    for (unsigned probe = 0; probe <= mask; ++probe, h = (h + 1) & mask) {
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&tab[h].key);
        if (cur == key) return h;
        if (cur == 0ull) {
            if (!create) return 0xFFFFFFFFu;
            cur = atomicCAS(&tab[h].key, 0ull, key);
            if (cur == 0ull || cur == key) return h;
        }
    }
    return 0xFFFFFFFFu;
}

// ---- update (addKeypoint for a whole frame) --------------------------------------------------------------------------
// world position, lattice snap, block slot of every keypoint; first toucher of a block lists it
__global__ void gm_prepare_kernel(const float4* __restrict__ kp, const float* __restrict__ ratio, const int* __restrict__ count_dev,
                                  unsigned n_cap, const float* __restrict__ pose, int have_pose, GmapBlock* tab, unsigned mask,
                                  float4* __restrict__ world, unsigned* __restrict__ blk_of, unsigned* __restrict__ touched_list,
                                  unsigned* __restrict__ ctl, unsigned epoch) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned n = count_dev ? min(n_cap, (unsigned)max(*count_dev, 0)) : n_cap;
    if (i >= n) return;
    const float4 p = kp[i];
    float x = p.x, y = p.y, z = p.z;
    if (have_pose) {  // kp_pos = R * kp_pos + T (src/lidar_odometry.cpp:351), row-major 3x4 [R|T]
        x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pose[0], p.x), __fmul_rn(pose[1], p.y)), __fmul_rn(pose[2], p.z)), pose[3]);
        y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pose[4], p.x), __fmul_rn(pose[5], p.y)), __fmul_rn(pose[6], p.z)), pose[7]);
        z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pose[8], p.x), __fmul_rn(pose[9], p.y)), __fmul_rn(pose[10], p.z)), pose[11]);
    }
    // Keypoint::createKeypoint: int(trunc(pos / prec)) * prec
    x = (float)((int)truncf(x / GM_LATTICE) * (int)GM_LATTICE);
    y = (float)((int)truncf(y / GM_LATTICE) * (int)GM_LATTICE);
    z = (float)((int)truncf(z / GM_LATTICE) * (int)GM_LATTICE);
    world[i] = make_float4(x, y, z, ratio[i]);
    const bool ok = isfinite(x) && isfinite(y) && isfinite(z);
    unsigned slot = 0xFFFFFFFFu;
    if (ok) slot = gm_find(tab, mask, gm_block_id(x, y, z), true);
    blk_of[i] = slot;
    if (slot == 0xFFFFFFFFu) { if (ok) atomicAdd(&ctl[3], 1u); return; }  // table full: counted, reported by the host
    if (atomicExch(&tab[slot].touched, epoch) != epoch) touched_list[atomicAdd(&ctl[1], 1u)] = slot;
}

// one warp per touched block: the frame's keypoints of that block, in keypoint order, through the admission rule
__global__ void __launch_bounds__(128)
gm_admit_kernel(const float4* __restrict__ world, const unsigned* __restrict__ blk_of, const uint64_t* __restrict__ bits,
                const int* __restrict__ count_dev, unsigned n_cap, GmapBlock* tab, const unsigned* __restrict__ touched_list,
                unsigned* __restrict__ ctl, float4* __restrict__ e_pos, uint64_t* __restrict__ e_desc, unsigned* __restrict__ chunks,
                unsigned max_entries, unsigned max_chunks) {
    const unsigned lane = threadIdx.x & 31, w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= ctl[1]) return;
    const unsigned slot = touched_list[w];
    const unsigned n = count_dev ? min(n_cap, (unsigned)max(*count_dev, 0)) : n_cap;
    unsigned cnt = tab[slot].count, first = tab[slot].first, last = tab[slot].last;
    for (unsigned base = 0; base < n; base += 32) {
        const unsigned i = base + lane;
        unsigned members = __ballot_sync(0xffffffffu, i < n && blk_of[i] == slot);
        while (members) {
            const unsigned k = base + (unsigned)(__ffs(members) - 1);
            members &= members - 1u;
            const float4 p = world[k];  // (x, y, z, seg-ratio) of the new keypoint
            // walk the stored entries of the block, 32 at a time
            bool reject = false;
            unsigned same = 0xFFFFFFFFu;
            unsigned c = first;
            for (unsigned j0 = 0; j0 < cnt; j0 += 32, c = chunks[c * 33u + 32u]) {
                const unsigned j = j0 + lane;
                bool rj = false;
                unsigned sm = 0xFFFFFFFFu;
                if (j < cnt) {
                    const unsigned e = chunks[c * 33u + lane];
                    const float4 q = e_pos[e];
                    const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
                    const float dist = sqrtf(dot3_rn(dx, dy, dz, dx, dy, dz));
                    rj = dist < GM_REJECT_MM && p.w <= q.w;   // src/mymap.cpp:17-18
                    if (q.x == p.x && q.y == p.y && q.z == p.z) sm = e;
                }
                reject = reject || __any_sync(0xffffffffu, rj);
                const unsigned who = __ballot_sync(0xffffffffu, sm != 0xFFFFFFFFu);
                if (who) same = __shfl_sync(0xffffffffu, sm, __ffs(who) - 1);
            }
            if (reject) continue;
            unsigned e = same;  // keypoints_[block][position] = keypoint: overwrite at an equal position, else insert
            if (e == 0xFFFFFFFFu) {
                if (lane == 0) e = atomicAdd(&ctl[0], 1u);
                e = __shfl_sync(0xffffffffu, e, 0);
                if (e >= max_entries) { if (lane == 0) { atomicAdd(&ctl[3], 1u); atomicSub(&ctl[0], 1u); } continue; }
                if ((cnt & 31u) == 0u) {  // new chunk
                    unsigned nc = 0;
                    if (lane == 0) nc = atomicAdd(&ctl[2], 1u);
                    nc = __shfl_sync(0xffffffffu, nc, 0);
                    if (nc >= max_chunks) { if (lane == 0) atomicAdd(&ctl[3], 1u); continue; }
                    if (lane == 0) {
                        chunks[nc * 33u + 32u] = 0xFFFFFFFFu;
                        if (cnt == 0u) first = nc; else chunks[last * 33u + 32u] = nc;
                    }
                    first = __shfl_sync(0xffffffffu, first, 0);
                    last = nc;
                }
                if (lane == 0) chunks[last * 33u + (cnt & 31u)] = e;
                ++cnt;
            }
            if (lane == 0) e_pos[e] = p;
            if (lane < 6) e_desc[(size_t)e * 6 + lane] = bits[(size_t)k * 6 + lane];
            __syncwarp();
            __threadfence_block();
        }
    }
    if (lane == 0) { tab[slot].count = cnt; tab[slot].first = first; tab[slot].last = last; }
}

// ---- gather (getKeypoints) ----------------------------------------------------------------------------------------------
struct GmapCube { int x0, y0, z0, nx, ny, nz; };

__host__ __device__ inline GmapCube gm_cube(float px, float py, float pz, float range) {
    GmapCube c;
    const float prec = (float)GM_PREC;
    c.x0 = (int)roundf((px - range) / prec); c.nx = (int)roundf((px + range) / prec) - c.x0 + 1;
    c.y0 = (int)roundf((py - range) / prec); c.ny = (int)roundf((py + range) / prec) - c.y0 + 1;
    c.z0 = (int)roundf((pz - range) / prec); c.nz = (int)roundf((pz + range) / prec) - c.z0 + 1;
    return c;
}

// one CTA: counts of the probed blocks in loop order (x outermost, z innermost), exclusive scan -> offsets
__global__ void __launch_bounds__(1024)
gm_probe_kernel(const GmapBlock* __restrict__ tab, unsigned mask, GmapCube cube, unsigned* __restrict__ probe_slot,
                unsigned* __restrict__ probe_off, unsigned* __restrict__ ctl, unsigned max_probe) {
    __shared__ unsigned ws[32];
    __shared__ unsigned carry;
    const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned np = min((unsigned)(cube.nx * cube.ny * cube.nz), max_probe);
    if (tid == 0) carry = 0;
    __syncthreads();
    for (unsigned base = 0; base < np; base += 1024) {
        const unsigned p = base + tid;
        unsigned cnt = 0, slot = 0xFFFFFFFFu;
        if (p < np) {
            const int iz = (int)(p % (unsigned)cube.nz), iy = (int)((p / (unsigned)cube.nz) % (unsigned)cube.ny), ix = (int)(p / ((unsigned)cube.nz * (unsigned)cube.ny));
            const float x = (float)((cube.x0 + ix) * (int)GM_PREC), y = (float)((cube.y0 + iy) * (int)GM_PREC), z = (float)((cube.z0 + iz) * (int)GM_PREC);
            slot = gm_find(const_cast<GmapBlock*>(tab), mask, gm_block_id(x, y, z), false);
            if (slot != 0xFFFFFFFFu) cnt = tab[slot].count;
            probe_slot[p] = slot;
        }
        unsigned inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += up;
        }
        if (lane == 31) ws[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const unsigned v = ws[lane];
            unsigned winc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned up = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= (unsigned)o) winc += up;
            }
            ws[lane] = winc - v;
        }
        __syncthreads();
        const unsigned excl = carry + ws[wid] + inc - cnt;
        if (p < np) probe_off[p] = excl;
        __syncthreads();
        if (tid == 1023) carry = excl + cnt;
        __syncthreads();
    }
    if (tid == 0) { ctl[4] = carry; ctl[5] = np; }
}

// one warp per probed block: entries in insertion order -> out[offset + j]
__global__ void __launch_bounds__(128)
gm_collect_kernel(const GmapBlock* __restrict__ tab, const unsigned* __restrict__ probe_slot, const unsigned* __restrict__ probe_off,
                  const unsigned* __restrict__ ctl, const unsigned* __restrict__ chunks, const float4* __restrict__ e_pos,
                  const uint64_t* __restrict__ e_desc, float4* __restrict__ out_pos, uint64_t* __restrict__ out_desc, unsigned out_cap) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned np = ctl[5];
    for (unsigned p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < np; p += (gridDim.x * blockDim.x) >> 5) {
        const unsigned slot = probe_slot[p];
        if (slot == 0xFFFFFFFFu) continue;
        const unsigned cnt = tab[slot].count, off = probe_off[p];
        unsigned c = tab[slot].first;
        for (unsigned j0 = 0; j0 < cnt; j0 += 32, c = chunks[c * 33u + 32u]) {
            const unsigned j = j0 + lane;
            if (j < cnt && off + j < out_cap) {
                const unsigned e = chunks[c * 33u + lane];
                out_pos[off + j] = e_pos[e];
#pragma unroll
                for (int k = 0; k < 6; ++k) out_desc[(size_t)(off + j) * 6 + k] = e_desc[(size_t)e * 6 + k];
            }
        }
    }
}

// appends the reference frame: positions transformed by its pose (pcl::transformPointCloud, :202), descriptors as they are
__global__ void gm_append_ref_kernel(const float4* __restrict__ ref_kp, const uint64_t* __restrict__ ref_bits, const int* __restrict__ ref_count,
                                     unsigned n_cap, const float* __restrict__ pose, int have_pose, const unsigned* __restrict__ ctl,
                                     float4* __restrict__ out_pos, uint64_t* __restrict__ out_desc, unsigned out_cap, unsigned* __restrict__ total_out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned n = ref_count ? min(n_cap, (unsigned)max(*ref_count, 0)) : n_cap;
    const unsigned base = min(ctl[4], out_cap);
    if (i == 0) *total_out = min(base + n, out_cap);
    if (i >= n || base + i >= out_cap) return;
    const float4 p = ref_kp[i];
    float x = p.x, y = p.y, z = p.z;
    if (have_pose) {
        x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pose[0], p.x), __fmul_rn(pose[1], p.y)), __fmul_rn(pose[2], p.z)), pose[3]);
        y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pose[4], p.x), __fmul_rn(pose[5], p.y)), __fmul_rn(pose[6], p.z)), pose[7]);
        z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pose[8], p.x), __fmul_rn(pose[9], p.y)), __fmul_rn(pose[10], p.z)), pose[11]);
    }
    out_pos[base + i] = make_float4(x, y, z, 0.0f);
#pragma unroll
    for (int k = 0; k < 6; ++k) out_desc[(size_t)(base + i) * 6 + k] = ref_bits[(size_t)i * 6 + k];
}

// ---- host side -----------------------------------------------------------------------------------------------------------
int gmap_create(Ctx* c, size_t max_entries, size_t max_blocks) {
    Gmap& g = c->gmap;
    if (g.d_tab) return BSHOT_OK;
    unsigned cap = 1024;
    while (cap < 2 * max_blocks) cap <<= 1;
    g.tab_cap = cap;
    g.max_entries = (unsigned)max_entries;
    g.max_chunks = (unsigned)(max_entries / 32 + max_blocks + 64);
    g.max_probe = 32768;
    BSHOT_CUDA_TRY(cudaMalloc((void**)&g.d_tab, sizeof(GmapBlock) * cap));
    BSHOT_CUDA_TRY(cudaMalloc((void**)&g.d_epos, sizeof(float4) * max_entries));
    BSHOT_CUDA_TRY(cudaMalloc((void**)&g.d_edesc, 48 * max_entries));
    BSHOT_CUDA_TRY(cudaMalloc((void**)&g.d_chunks, sizeof(unsigned) * 33 * g.max_chunks));
    BSHOT_CUDA_TRY(cudaMalloc((void**)&g.d_world, sizeof(float4) * c->max_kp));
    BSHOT_CUDA_TRY(cudaMalloc((void**)&g.d_blk_of, sizeof(unsigned) * c->max_kp));
    BSHOT_CUDA_TRY(cudaMalloc((void**)&g.d_touched, sizeof(unsigned) * c->max_kp));
    BSHOT_CUDA_TRY(cudaMalloc((void**)&g.d_probe_slot, sizeof(unsigned) * g.max_probe));
    BSHOT_CUDA_TRY(cudaMalloc((void**)&g.d_probe_off, sizeof(unsigned) * g.max_probe));
    BSHOT_CUDA_TRY(cudaMalloc((void**)&g.d_ctl, sizeof(unsigned) * 8));
    BSHOT_CUDA_TRY(cudaMalloc((void**)&g.d_pose, sizeof(float) * 24));
    BSHOT_CUDA_TRY(cudaMalloc((void**)&g.d_tpos, sizeof(float4) * c->max_targets));
    return gmap_reset(c);
}

void gmap_free(Ctx* c) {
    Gmap& g = c->gmap;
    void* ptrs[] = {g.d_tab, g.d_epos, g.d_edesc, g.d_chunks, g.d_world, g.d_blk_of, g.d_touched, g.d_probe_slot, g.d_probe_off, g.d_ctl, g.d_pose, g.d_tpos};
    for (void* p : ptrs) if (p) cudaFree(p);
    g = Gmap();
}

int gmap_reset(Ctx* c) {
    Gmap& g = c->gmap;
    if (!g.d_tab) return BSHOT_OK;
    BSHOT_CUDA_TRY(cudaMemsetAsync(g.d_tab, 0, sizeof(GmapBlock) * g.tab_cap, c->stream));
    BSHOT_CUDA_TRY(cudaMemsetAsync(g.d_ctl, 0, sizeof(unsigned) * 8, c->stream));
    g.epoch = 0;
    g.n_gathered = 0;
    return BSHOT_OK;
}

// ctl: [0] entries, [1] touched blocks of the current update, [2] chunks, [3] dropped (capacity), [4] gathered map entries, [5] probes
int gmap_update(Ctx* c, const float4* d_kp, const float* d_ratio, const uint64_t* d_bits, const int* d_count, size_t n_cap, const float* pose12) {
    Gmap& g = c->gmap;
    if (!g.d_tab) { set_error("map: bshot_gmap_create first"); return BSHOT_E_STATE; }
    if (n_cap == 0) return BSHOT_OK;
    if (n_cap > c->max_kp) { set_error("map update: %zu keypoints > capacity %zu", n_cap, c->max_kp); return BSHOT_E_CAPACITY; }
    if (pose12) BSHOT_CUDA_TRY(cudaMemcpyAsync(g.d_pose, pose12, sizeof(float) * 12, cudaMemcpyHostToDevice, c->stream));
    BSHOT_CUDA_TRY(cudaMemsetAsync(g.d_ctl + 1, 0, sizeof(unsigned), c->stream));
    ++g.epoch;
    const unsigned n = (unsigned)n_cap;
    gm_prepare_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_kp, d_ratio, d_count, n, g.d_pose, pose12 ? 1 : 0, reinterpret_cast<GmapBlock*>(g.d_tab),
                                                             g.tab_cap - 1, g.d_world, g.d_blk_of, g.d_touched, g.d_ctl, g.epoch);
    // at most n blocks are touched: one warp each (warps beyond the device-side count exit)
    gm_admit_kernel<<<(n * 32 + 127) / 128, 128, 0, c->stream>>>(g.d_world, g.d_blk_of, d_bits, d_count, n, reinterpret_cast<GmapBlock*>(g.d_tab), g.d_touched,
                                                                g.d_ctl, g.d_epos, g.d_edesc, g.d_chunks, g.max_entries, g.max_chunks);
    count_launch(c, 2);
    return check_launch("map update kernels");
}

// target set of featureMatching in RUN status: map keypoints within `range` of `pos`, then the reference frame
// (d_ref_* may be null).  Descriptors -> d_t_out (the matcher's target buffer), positions -> g.d_tpos; *d_total = count.
int gmap_gather(Ctx* c, const float pos[3], float range, const float4* d_ref_kp, const uint64_t* d_ref_bits, const int* d_ref_count,
                size_t ref_cap, const float* ref_pose12, uint64_t* d_t_out, size_t out_cap, unsigned* d_total) {
    Gmap& g = c->gmap;
    if (!g.d_tab) { set_error("map: bshot_gmap_create first"); return BSHOT_E_STATE; }
    const GmapCube cube = gm_cube(pos[0], pos[1], pos[2], range);
    if ((long long)cube.nx * cube.ny * cube.nz > (long long)g.max_probe) { set_error("map gather: range %.0f probes more than %u blocks", range, g.max_probe); return BSHOT_E_CAPACITY; }
    if (ref_pose12) BSHOT_CUDA_TRY(cudaMemcpyAsync(g.d_pose + 12, ref_pose12, sizeof(float) * 12, cudaMemcpyHostToDevice, c->stream));
    gm_probe_kernel<<<1, 1024, 0, c->stream>>>(reinterpret_cast<const GmapBlock*>(g.d_tab), g.tab_cap - 1, cube, g.d_probe_slot, g.d_probe_off, g.d_ctl, g.max_probe);
    gm_collect_kernel<<<(unsigned)c->sm_count * 2u, 128, 0, c->stream>>>(reinterpret_cast<const GmapBlock*>(g.d_tab), g.d_probe_slot, g.d_probe_off, g.d_ctl, g.d_chunks,
                                                                        g.d_epos, g.d_edesc, g.d_tpos, d_t_out, (unsigned)out_cap);
    const unsigned nref = d_ref_bits ? (unsigned)ref_cap : 0u;
    gm_append_ref_kernel<<<(std::max(nref, 1u) + 255) / 256, 256, 0, c->stream>>>(d_ref_kp, d_ref_bits, d_ref_count, nref, g.d_pose + 12, ref_pose12 ? 1 : 0, g.d_ctl,
                                                                                g.d_tpos, d_t_out, (unsigned)out_cap, d_total);
    count_launch(c, 3);
    return check_launch("map gather kernels");
}

}  // namespace bshot

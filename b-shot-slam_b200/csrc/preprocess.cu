// preprocess.cu -- the scan preprocessor on the device (SURVEY 8f "next" #4): one rotation of raw laser returns -> the
// cloud the front end works on.  Replaces myslam::Preprocessor::run (src/preprocess.cpp:213-223): readFrame (:38-71: range
// image keyed by azimuth / vertical angle, a synthetic start entry below the lowest beam), removeGround (:73-166: a state
// machine up every azimuth column), removeOccluded (:168-197: range jumps between neighbouring columns of one ring mark the
// far side), writePointCloud (:199-211: kept points in azimuth-major, vertical-ascending order).  The reference keeps the
// range image in std::map<double, std::map<double, double>>; here the returns are cut into runs of equal azimuth (a firing),
// the runs are sorted by azimuth (stable radix sort, one CTA: `capture >> lasers` delivers a rotation in firing order, which
// wraps through 0 degrees somewhere), equal azimuths merge into one column and a thread sorts its <= 128 returns by vertical angle.
//   runs     flag + scan, 16 x 4-bit stable counting sort of the run keys, column heads + scan
//   columns  thread per column : sort + de-duplicate (last return wins), ground state machine, ring table
//   rings    warp per ring     : previous non-lost column by a max-scan over the columns, occlusion marks
//   output   thread per column : count kept, CTA-wide scan, write
// Every float / double conversion follows the reference's expression types (Vector3f holds floats; asin of a float
// argument is the float overload).  tests/test_preprocess.py checks it against the reference source compiled unchanged.
#include <math.h>
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "stages.h"

namespace bshot {

constexpr int PRE_LMAX = 128;   // returns per azimuth column (HDL-64E: 64)
constexpr double PRE_PI = 3.1415926535897932384626433832795;  // CV_PI

struct PreParams {
    double vert_init, lowpt_th;
    double grad_th, height_th, dist_th, angdiff_th;
};

// column starts: flag where the azimuth changes, exclusive scan by one CTA (also used for the output offsets)
__global__ void __launch_bounds__(1024)
pre_scan_kernel(const unsigned* __restrict__ in, unsigned n, const unsigned* __restrict__ n_dev, unsigned* __restrict__ out_excl, unsigned* __restrict__ total) {
    __shared__ unsigned ws[32];
    __shared__ unsigned carry;
    if (n_dev) n = *n_dev;
    const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (unsigned base = 0; base < n; base += 1024) {
        const unsigned i = base + tid;
        const unsigned v = (i < n) ? in[i] : 0u;
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += up;
        }
        if (lane == 31) ws[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const unsigned w = ws[lane];
            unsigned winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned up = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= (unsigned)o) winc += up;
            }
            ws[lane] = winc - w;
        }
        __syncthreads();
        const unsigned excl = carry + ws[wid] + inc - v;
        if (i < n) out_excl[i] = excl;
        __syncthreads();
        if (tid == 1023) carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) *total = carry;
}

__device__ __forceinline__ unsigned long long orderable(double v) {
    v += 0.0;  // -0.0 and +0.0 are one std::map key
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

// azimuth key of every return (radians, readFrame :47) and the heads of the runs of equal keys
__global__ void pre_flag_kernel(const double* __restrict__ az_deg, unsigned n, double* __restrict__ azr, unsigned* __restrict__ flag) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = az_deg[i] * PRE_PI / 180.0;
    azr[i] = a;
    flag[i] = (i == 0 || a != az_deg[i - 1] * PRE_PI / 180.0) ? 1u : 0u;
}

__global__ void pre_runstart_kernel(const unsigned* __restrict__ flag, const unsigned* __restrict__ excl, const double* __restrict__ azr, unsigned n,
                                    unsigned* __restrict__ run_start, unsigned long long* __restrict__ run_key, unsigned* __restrict__ run_idx) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) {
        const unsigned r = excl[i];
        run_start[r] = i;
        run_key[r] = orderable(azr[i]);
        run_idx[r] = r;
    }
}

// stable LSD radix sort of the runs by key, 4 bits a pass, one CTA: every thread owns a contiguous chunk, so the order of
// equal keys (= input order, "a later return wins") survives.  16 passes: the result is back in (key_a, idx_a).
__global__ void __launch_bounds__(1024)
pre_sort_runs_kernel(unsigned long long* __restrict__ key_a, unsigned* __restrict__ idx_a, unsigned long long* __restrict__ key_b, unsigned* __restrict__ idx_b,
                     const unsigned* __restrict__ nrun_dev) {
    extern __shared__ unsigned cnt[];   // 16 x 1024 counters
    __shared__ unsigned ws[32];
    __shared__ unsigned carry;
    const unsigned R = *nrun_dev, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned chunk = (R + 1023) / 1024, s = min(R, tid * chunk), e = min(R, s + chunk);
    unsigned long long* ka = key_a; unsigned* ia = idx_a; unsigned long long* kb = key_b; unsigned* ib = idx_b;
    for (int pass = 0; pass < 16; ++pass) {
        const int sh = 4 * pass;
        unsigned c[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) c[d] = 0;
        for (unsigned i = s; i < e; ++i) {
            const unsigned d = (unsigned)(ka[i] >> sh) & 15u;
#pragma unroll
            for (int k = 0; k < 16; ++k) c[k] += (d == (unsigned)k);
        }
#pragma unroll
        for (int d = 0; d < 16; ++d) cnt[d * 1024 + tid] = c[d];
        if (tid == 0) carry = 0;
        __syncthreads();
        // exclusive scan over the 16 x 1024 counters in digit-major order
        for (int d = 0; d < 16; ++d) {
            const unsigned v = cnt[d * 1024 + tid];
            unsigned inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned up = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= (unsigned)o) inc += up;
            }
            if (lane == 31) ws[wid] = inc;
            __syncthreads();
            if (wid == 0) {
                const unsigned w = ws[lane];
                unsigned winc = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned up = __shfl_up_sync(0xffffffffu, winc, o);
                    if (lane >= (unsigned)o) winc += up;
                }
                ws[lane] = winc - w;
            }
            __syncthreads();
            const unsigned excl = carry + ws[wid] + inc - v;
            cnt[d * 1024 + tid] = excl;
            __syncthreads();
            if (tid == 1023) carry = excl + v;
            __syncthreads();
        }
#pragma unroll
        for (int d = 0; d < 16; ++d) c[d] = cnt[d * 1024 + tid];
        for (unsigned i = s; i < e; ++i) {
            const unsigned long long k = ka[i];
            const unsigned d = (unsigned)(k >> sh) & 15u;
            unsigned o = 0;
#pragma unroll
            for (int q = 0; q < 16; ++q) if (d == (unsigned)q) { o = c[q]; c[q] = o + 1; }
            kb[o] = k; ib[o] = ia[i];
        }
        __syncthreads();
        unsigned long long* tk = ka; ka = kb; kb = tk;
        unsigned* ti = ia; ia = ib; ib = ti;
    }
}

// heads of the columns: sorted runs whose key differs from the one before
__global__ void pre_colflag_kernel(const unsigned long long* __restrict__ key, const unsigned* __restrict__ nrun_dev, unsigned* __restrict__ flag) {
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= *nrun_dev) return;
    flag[j] = (j == 0 || key[j] != key[j - 1]) ? 1u : 0u;
}

__global__ void pre_colstart_kernel(const unsigned* __restrict__ flag, const unsigned* __restrict__ excl, const unsigned* __restrict__ nrun_dev,
                                    unsigned* __restrict__ col_first) {
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < *nrun_dev && flag[j]) col_first[excl[j]] = j;
}

// Eigen's float norm() = sqrt(squaredNorm()), summed in x, y, z order with no fused multiply-add (the reference is built
// for baseline x86-64)
__device__ __forceinline__ float norm3(float x, float y, float z) {
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

// per column: sorted (vertical, distance) entries incl. the synthetic start entry, ground / self-car flags, ring table
__global__ void __launch_bounds__(64)
pre_columns_kernel(const double* __restrict__ azr, const double* __restrict__ vert_deg, const unsigned short* __restrict__ dist_u16,
                   const unsigned char* __restrict__ sel_in, unsigned n,
                   const unsigned* __restrict__ run_start, const unsigned* __restrict__ run_idx, const unsigned* __restrict__ nrun_dev,
                   const unsigned* __restrict__ col_first, const unsigned* __restrict__ ncol_dev, const double* __restrict__ ring_rad, unsigned nv,
                   PreParams P, double* __restrict__ c_vert, double* __restrict__ c_dist, unsigned char* __restrict__ c_rm, unsigned char* __restrict__ c_sel,
                   unsigned* __restrict__ c_cnt, double* __restrict__ c_az, double* __restrict__ ring_dist, int* __restrict__ ring_ent, unsigned* __restrict__ err) {
    const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned ncol = *ncol_dev, nrun = *nrun_dev;
    if (c >= ncol) return;
    const unsigned j0 = col_first[c], j1 = (c + 1 < ncol) ? col_first[c + 1] : nrun;
    double vk[PRE_LMAX + 1], dk[PRE_LMAX + 1];
    unsigned char rm[PRE_LMAX + 1], sl[PRE_LMAX + 1];   // sl: selmap (readFrame :58-67), a later return overwrites it too
    int m = 0;
    const double az = azr[run_start[run_idx[j0]]];
    // rimg[azimuth][vertical] = distance: std::map semantics = sorted by key, a later return with the same key wins
    for (unsigned j = j0; j < j1; ++j) {
        const unsigned r = run_idx[j], s = run_start[r], e = (r + 1 < nrun) ? run_start[r + 1] : n;
        for (unsigned i = s; i < e; ++i) {
            const double v = vert_deg[i] * PRE_PI / 180.0, d = (double)dist_u16[i] * 2.0;   // :46,:48
            int pos = 0;
            while (pos < m && vk[pos] < v) ++pos;
            const unsigned char sv = sel_in ? sel_in[i] : (unsigned char)1;
            if (pos < m && vk[pos] == v) { dk[pos] = d; sl[pos] = sv; continue; }
            if (m >= PRE_LMAX) { *err = 2u; return; }
            for (int q = m; q > pos; --q) { vk[q] = vk[q - 1]; dk[q] = dk[q - 1]; sl[q] = sl[q - 1]; }
            vk[pos] = v; dk[pos] = d; sl[pos] = sv; ++m;
        }
    }
    {   // rimg[azimuth][vert_init_] = 2450 / sin(vert_init_), rmmap = 1 (:55-57), written last: it wins over an equal key
        const double v = P.vert_init, d = 2450.0 / sin(P.vert_init);
        int pos = 0;
        while (pos < m && vk[pos] < v) ++pos;
        if (pos < m && vk[pos] == v) dk[pos] = d;
        else {
            for (int j = m; j > pos; --j) { vk[j] = vk[j - 1]; dk[j] = dk[j - 1]; sl[j] = sl[j - 1]; }
            vk[pos] = v; dk[pos] = d; sl[pos] = 0; ++m;
        }
    }
    for (int j = 0; j < m; ++j) rm[j] = (vk[j] == P.vert_init) ? 1 : 0;
    // ---- removeGround (:73-166) up the column ---------------------------------------------------------------------------
    {
        bool lost_pt = false, set_th_pt = false, prev_is_ground = true;
        const double x_0 = (-2450.0 / tan(P.vert_init)) * sin(az), y_0 = (-2450.0 / tan(P.vert_init)) * cos(az), z_0 = -2450.0;
        float pp[3] = {(float)x_0, (float)y_0, (float)z_0}, pth[3] = {(float)x_0, (float)y_0, (float)z_0};
        for (int j = 1; j < m; ++j) {  // the first entry of the column is skipped (:89-92)
            const double vert = vk[j], dist = dk[j];
            const double x = dist * cos(vert) * sin(az), y = dist * cos(vert) * cos(az), z = dist * sin(vert);
            const float pc[3] = {(float)x, (float)y, (float)z};
            const float dx = pc[0] - pp[0], dy = pc[1] - pp[1], dz = pc[2] - pp[2];
            const float dn = norm3(dx, dy, dz);
            // asin((float) / (float)) is the float overload; * 180 stays float; / CV_PI promotes to double
            const float as = (float)asin((double)(dz / dn));
            const double grad = (double)(as * 180.0f) / PRE_PI;
            const float pp_norm = norm3(pp[0], pp[1], pp[2]);
            unsigned char r = rm[j];
            if (prev_is_ground && (grad > P.grad_th || dist == 0.0 || dist < (double)pp_norm)) {
                set_th_pt = true;
                pth[0] = pp[0]; pth[1] = pp[1]; pth[2] = pp[2];
            }
            if (prev_is_ground) {
                if (grad < P.grad_th && !lost_pt) { r = 1; prev_is_ground = true; }
                else { r = 0; prev_is_ground = false; }
            } else if (!prev_is_ground && (double)pc[2] < P.lowpt_th && grad < P.grad_th) {
                r = 1; prev_is_ground = true; set_th_pt = false;
            }
            if (dist == 0.0) { r = 1; lost_pt = true; prev_is_ground = false; }
            else lost_pt = false;
            if (dist < (double)pp_norm && dist != 0.0) { r = 0; prev_is_ground = false; }
            if (set_th_pt && (double)(pc[2] - pth[2]) < P.height_th && pc[2] < pp[2]) { set_th_pt = false; r = 1; prev_is_ground = true; }
            if (x <= 820 && x >= -820 && y <= 1300 && y >= -1800 && z <= 100 && z >= -2000) r = 2;  // self-car
            rm[j] = r;
            pp[0] = pc[0]; pp[1] = pc[1]; pp[2] = pc[2];
        }
    }
    const size_t base = (size_t)c * (PRE_LMAX + 1);
    for (int j = 0; j < m; ++j) { c_vert[base + j] = vk[j]; c_dist[base + j] = dk[j]; c_rm[base + j] = rm[j]; c_sel[base + j] = sl[j]; }
    c_cnt[c] = (unsigned)m;
    c_az[c] = az;
    // ring table for removeOccluded: rimg[col][v] of every ring key (0 when the column has no such entry)
    for (unsigned r = 0; r < nv; ++r) {
        const double v = ring_rad[r];
        int lo = 0, hi = m;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (vk[mid] < v) lo = mid + 1; else hi = mid; }
        const bool hit = lo < m && vk[lo] == v;
        ring_dist[(size_t)r * ncol + c] = hit ? dk[lo] : 0.0;
        ring_ent[(size_t)r * ncol + c] = hit ? lo : -1;
    }
}

// removeOccluded (:168-197): one warp per ring walks the columns 32 at a time; prev_hor of a column = the last earlier
// column that was the first one or held a return on this ring
__global__ void __launch_bounds__(128)
pre_rings_kernel(const double* __restrict__ ring_dist, const int* __restrict__ ring_ent, const double* __restrict__ c_az, const unsigned* __restrict__ ncol_dev,
                 unsigned nv, PreParams P, unsigned char* __restrict__ c_rm) {
    const unsigned lane = threadIdx.x & 31, r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= nv) return;
    const unsigned ncol = *ncol_dev;
    const double* dist = ring_dist + (size_t)r * ncol;
    const int* ent = ring_ent + (size_t)r * ncol;
    int carry = 0;  // prev_hor before the current batch (column 0 to start with)
    for (unsigned base = 0; base < ncol; base += 32) {
        const unsigned c = base + lane;
        const double d = (c < ncol) ? dist[c] : 0.0;
        const bool updates = (c < ncol) && (c == 0 || d != 0.0);   // this column becomes prev_hor for the ones after it
        int mine = updates ? (int)c : -1;
        int inc = mine;   // inclusive max-scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc = max(inc, up);
        }
        int prev = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) prev = -1;
        prev = max(prev, carry);
        if (c < ncol && c != 0 && d != 0.0) {
            const double d_dist = d - dist[prev], d_hor = c_az[c] - c_az[prev];
            if (fabs(d_dist) > P.dist_th && fabs(d_hor) < P.angdiff_th) {
                if (d_dist > 0) {  // the current point is the background
                    unsigned char* p = c_rm + (size_t)c * (PRE_LMAX + 1) + ent[c];
                    if (*p == 0) *p = 3;
                } else if (ent[prev] >= 0) {  // the previous one is
                    unsigned char* p = c_rm + (size_t)prev * (PRE_LMAX + 1) + ent[prev];
                    if (*p == 0) *p = 3;
                }
            }
        }
        carry = max(carry, __shfl_sync(0xffffffffu, inc, 31));
    }
}

__global__ void pre_count_kernel(const double* __restrict__ c_vert, const double* __restrict__ c_dist, const unsigned char* __restrict__ c_rm,
                                 const unsigned char* __restrict__ c_sel, unsigned char save_sel, const unsigned* __restrict__ c_cnt,
                                 const unsigned* __restrict__ ncol_dev, double vert_init, unsigned* __restrict__ keep) {
    const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= *ncol_dev) return;
    const size_t base = (size_t)c * (PRE_LMAX + 1);
    unsigned k = 0;
    for (unsigned j = 0; j < c_cnt[c]; ++j)
        if (c_dist[base + j] != 0.0 && c_vert[base + j] != vert_init && c_rm[base + j] == 0 && c_sel[base + j] == save_sel) ++k;   // :201-209
    keep[c] = k;
}

__global__ void pre_write_kernel(const double* __restrict__ c_vert, const double* __restrict__ c_dist, const unsigned char* __restrict__ c_rm,
                                 const unsigned char* __restrict__ c_sel, unsigned char save_sel, const unsigned* __restrict__ c_cnt, const double* __restrict__ c_az, const unsigned* __restrict__ ncol_dev, double vert_init,
                                 const unsigned* __restrict__ off, float* __restrict__ xyz, unsigned cap) {
    const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= *ncol_dev) return;
    const size_t base = (size_t)c * (PRE_LMAX + 1);
    const double az = c_az[c];
    unsigned o = off[c];
    for (unsigned j = 0; j < c_cnt[c]; ++j) {
        const double dist = c_dist[base + j], vert = c_vert[base + j];
        if (dist == 0.0 || vert == vert_init || c_rm[base + j] != 0 || c_sel[base + j] != save_sel) continue;
        if (o < cap) {
            xyz[3 * (size_t)o] = (float)(dist * cos(vert) * sin(az));
            xyz[3 * (size_t)o + 1] = (float)(dist * cos(vert) * cos(az));
            xyz[3 * (size_t)o + 2] = (float)(dist * sin(vert));
        }
        ++o;
    }
}

int scratch_reserve(Ctx* c, int which, size_t bytes) {
    if (c->pre_bytes[which] >= bytes) return BSHOT_OK;
    if (c->d_pre[which]) { cudaFree(c->d_pre[which]); c->d_pre[which] = nullptr; c->pre_bytes[which] = 0; }
    bytes += bytes / 4;
    BSHOT_CUDA_TRY(cudaMalloc(&c->d_pre[which], bytes));
    c->pre_bytes[which] = bytes;
    return BSHOT_OK;
}

struct Carver {
    char* p; size_t off = 0;
    template <typename T> T* take(size_t count) {
        T* r = reinterpret_cast<T*>(p + off);
        off += (count * sizeof(T) + 255) & ~(size_t)255;
        return r;
    }
};

// host inputs -> host output (xyz_out holds cap points).  Synchronous: two stream synchronisations (the number of columns
// sizes the per-column scratch; the number of kept points sizes the copy back).
int preprocess_run(Ctx* c, const double* az_deg, const double* vert_deg, const unsigned short* dist, size_t n, const double* ring_deg, size_t nv,
                   double vert_init, double lowpt_th, const unsigned char* sel, int save_sel, float* xyz_out, size_t cap, size_t* n_out,
                   const float** d_xyz_out) {
    if (d_xyz_out) *d_xyz_out = nullptr;
    if (n_out) *n_out = 0;
    if (n == 0) return BSHOT_OK;
    if (nv > 256) { set_error("bshot_preprocess: more than 256 rings"); return BSHOT_E_INVALID; }
    PreParams P;
    P.vert_init = vert_init; P.lowpt_th = lowpt_th;
    P.grad_th = 45; P.height_th = 500; P.dist_th = 3000; P.angdiff_th = 1.0 * PRE_PI / 180.0;   // include/preprocess.h:43-47
    std::vector<double> ring(std::max<size_t>(nv, 1), 0.0);
    for (size_t i = 0; i < nv; ++i) ring[i] = ring_deg[i] * PRE_PI / 180.0;   // removeOccluded :171
    const unsigned nn = (unsigned)n;
    const size_t pad = 256 * 16;
    BSHOT_TRY(scratch_reserve(c, 0, n * (8 + 8 + 2 + 1 + 8 + 4 + 4 + 4 + 8 + 4 + 8 + 4 + 4 + 4 + 4) + 8 * 256 + 12 * cap + pad));
    Carver A{(char*)c->d_pre[0]};
    double* d_az = A.take<double>(n); double* d_vert = A.take<double>(n); unsigned short* d_dist = A.take<unsigned short>(n);
    unsigned char* d_sel = A.take<unsigned char>(n);
    double* d_azr = A.take<double>(n);
    unsigned* d_flag = A.take<unsigned>(n); unsigned* d_excl = A.take<unsigned>(n); unsigned* d_run_start = A.take<unsigned>(n);
    unsigned long long* d_key_a = A.take<unsigned long long>(n); unsigned* d_idx_a = A.take<unsigned>(n);
    unsigned long long* d_key_b = A.take<unsigned long long>(n); unsigned* d_idx_b = A.take<unsigned>(n);
    unsigned* d_col_first = A.take<unsigned>(n);
    double* d_ring = A.take<double>(256);
    unsigned* ctl = A.take<unsigned>(8);   // [0] runs [1] error [2] kept [3] columns
    float* d_out = A.take<float>(3 * std::max<size_t>(cap, 1));
    static bool attr_set = false;
    if (!attr_set) {
        BSHOT_CUDA_TRY(cudaFuncSetAttribute(pre_sort_runs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024 * 4));
        attr_set = true;
    }
    BSHOT_CUDA_TRY(cudaMemcpyAsync(d_az, az_deg, 8 * n, cudaMemcpyHostToDevice, c->stream));
    BSHOT_CUDA_TRY(cudaMemcpyAsync(d_vert, vert_deg, 8 * n, cudaMemcpyHostToDevice, c->stream));
    BSHOT_CUDA_TRY(cudaMemcpyAsync(d_dist, dist, 2 * n, cudaMemcpyHostToDevice, c->stream));
    if (sel) BSHOT_CUDA_TRY(cudaMemcpyAsync(d_sel, sel, n, cudaMemcpyHostToDevice, c->stream));
    BSHOT_CUDA_TRY(cudaMemcpyAsync(d_ring, ring.data(), 8 * ring.size(), cudaMemcpyHostToDevice, c->stream));
    BSHOT_CUDA_TRY(cudaMemsetAsync(ctl, 0, 32, c->stream));
    const unsigned gb = (nn + 255) / 256;
    pre_flag_kernel<<<gb, 256, 0, c->stream>>>(d_az, nn, d_azr, d_flag);
    pre_scan_kernel<<<1, 1024, 0, c->stream>>>(d_flag, nn, nullptr, d_excl, ctl);
    pre_runstart_kernel<<<gb, 256, 0, c->stream>>>(d_flag, d_excl, d_azr, nn, d_run_start, d_key_a, d_idx_a);
    pre_sort_runs_kernel<<<1, 1024, 16 * 1024 * 4, c->stream>>>(d_key_a, d_idx_a, d_key_b, d_idx_b, ctl);
    pre_colflag_kernel<<<gb, 256, 0, c->stream>>>(d_key_a, ctl, d_flag);
    pre_scan_kernel<<<1, 1024, 0, c->stream>>>(d_flag, 0, ctl, d_excl, ctl + 3);
    pre_colstart_kernel<<<gb, 256, 0, c->stream>>>(d_flag, d_excl, ctl, d_col_first);
    count_launch(c, 7);
    BSHOT_TRY(check_launch("preprocess run kernels"));
    unsigned h_ctl[8];
    BSHOT_CUDA_TRY(cudaMemcpyAsync(h_ctl, ctl, 32, cudaMemcpyDeviceToHost, c->stream));
    BSHOT_CUDA_TRY(cudaStreamSynchronize(c->stream));
    const unsigned ncol = h_ctl[3];
    const size_t cw = (size_t)ncol * (PRE_LMAX + 1), rw = std::max<size_t>(nv, 1) * ncol;
    BSHOT_TRY(scratch_reserve(c, 1, cw * (8 + 8 + 1 + 1) + (size_t)ncol * (4 + 8 + 4 + 4) + rw * (8 + 4) + pad));
    Carver B{(char*)c->d_pre[1]};
    double* c_vert = B.take<double>(cw); double* c_dist = B.take<double>(cw); unsigned char* c_rm = B.take<unsigned char>(cw); unsigned char* c_sel = B.take<unsigned char>(cw);
    unsigned* c_cnt = B.take<unsigned>(ncol); double* c_az = B.take<double>(ncol);
    unsigned* d_keep = B.take<unsigned>(ncol); unsigned* d_off = B.take<unsigned>(ncol);
    double* d_rd = B.take<double>(rw); int* d_re = B.take<int>(rw);
    const unsigned* ncol_dev = ctl + 3;
    pre_columns_kernel<<<(ncol + 63) / 64, 64, 0, c->stream>>>(d_azr, d_vert, d_dist, sel ? d_sel : nullptr, nn, d_run_start, d_idx_a, ctl, d_col_first, ncol_dev, d_ring, (unsigned)nv, P,
                                                             c_vert, c_dist, c_rm, c_sel, c_cnt, c_az, d_rd, d_re, ctl + 1);
    if (nv) pre_rings_kernel<<<(unsigned)((nv * 32 + 127) / 128), 128, 0, c->stream>>>(d_rd, d_re, c_az, ncol_dev, (unsigned)nv, P, c_rm);
    pre_count_kernel<<<(ncol + 255) / 256, 256, 0, c->stream>>>(c_vert, c_dist, c_rm, c_sel, (unsigned char)(save_sel ? 1 : 0), c_cnt, ncol_dev, vert_init, d_keep);
    pre_scan_kernel<<<1, 1024, 0, c->stream>>>(d_keep, ncol, nullptr, d_off, ctl + 2);
    pre_write_kernel<<<(ncol + 255) / 256, 256, 0, c->stream>>>(c_vert, c_dist, c_rm, c_sel, (unsigned char)(save_sel ? 1 : 0), c_cnt, c_az, ncol_dev, vert_init, d_off, d_out, (unsigned)cap);
    count_launch(c, nv ? 5 : 4);
    BSHOT_TRY(check_launch("preprocess kernels"));
    BSHOT_CUDA_TRY(cudaMemcpyAsync(h_ctl, ctl, 32, cudaMemcpyDeviceToHost, c->stream));
    BSHOT_CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (h_ctl[1] == 2) { set_error("bshot_preprocess: more than %d returns in one azimuth column", PRE_LMAX); return BSHOT_E_CAPACITY; }
    const size_t kept = h_ctl[2];
    if (n_out) *n_out = kept;
    if (d_xyz_out) *d_xyz_out = d_out;   // valid until the next preprocessor call on this context
    if (xyz_out && kept) {
        BSHOT_CUDA_TRY(cudaMemcpyAsync(xyz_out, d_out, 12 * std::min(kept, cap), cudaMemcpyDeviceToHost, c->stream));
        BSHOT_CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    return BSHOT_OK;
}

}  // namespace bshot

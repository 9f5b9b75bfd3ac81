// ransac.cu -- RANSAC correspondence rejection on the device (SURVEY 8f "next" #1).
//
// Replaces the pcl::registration::CorrespondenceRejectorSampleConsensus call of LidarOdometry::featureMatching
// (src/lidar_odometry.cpp:251-261: 2000 iterations, inlier threshold 1500 mm) -- i.e. PCL 1.8's
// RandomSampleConsensus over SampleConsensusModelRegistration:
//   * sample sequence: mt19937 seeded 12345, rnd() = mt() / 2 (boost::uniform_int<>(0, INT_MAX)), drawIndexSample's
//     progressive shuffle of the index list, samples rejected while two of the three source points are closer than
//     sample_dist_thresh (mean of the square roots of the eigenvalues of the source covariance, squared) -- up to 1000 tries;
//   * model: Umeyama without scaling on the three pairs in double (SVD of the 3x3 cross-covariance, proper rotation),
//     coefficients cast to float;
//   * score: number of correspondences with |T * src - tgt|^2 < threshold^2 (float transform, float norm);
//   * adaptive stop: k = log(1 - 0.99) / log(1 - w^3) re-evaluated at every new best, first best wins ties;
//   * result: the correspondences within the threshold of the best model, in their original order, and that model.
// The sample sequence does not depend on the scores, so it is generated up front (host, a few microseconds), ALL
// max_iterations + 1 hypotheses are scored in parallel (one warp each), and the adaptive loop is replayed over the
// scores on the host exactly as PCL runs it.  This file is compiled with -fmad=false: the double-precision Jacobi SVD
// then uses only IEEE +, -, *, /, sqrt and is bit-identical to the oracle's restatement (orc_ransac).
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "rigid_math.cuh"
#include "stages.h"

namespace bshot {

// Umeyama without scaling for three point pairs (pcl::umeyama / Eigen::umeyama, with_scaling = false), double in, float
// 4x4 row-major out.  rank 2 (three non-collinear points): the third columns of U and V are completed as cross products,
// so both bases are right handed and R = U V^T is the proper rotation that Eq. (40)-(43) select.
__host__ __device__ inline void umeyama3(const double src[3][3], const double dst[3][3], float T[16]) {
    double sm[3], dm[3];
    for (int c = 0; c < 3; ++c) { sm[c] = (src[0][c] + src[1][c] + src[2][c]) / 3.0; dm[c] = (dst[0][c] + dst[1][c] + dst[2][c]) / 3.0; }
    double sigma[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double acc = 0.0;
            for (int i = 0; i < 3; ++i) acc += (dst[i][r] - dm[r]) * (src[i][c] - sm[c]);
            sigma[3 * r + c] = acc / 3.0;
        }
    double U[9], s[3], V[9];
    svd3_hestenes(sigma, U, s, V);
    // complete the last columns to right-handed bases (rank <= 2 for three points)
    U[2] = U[3] * U[7] - U[6] * U[4]; U[5] = U[6] * U[1] - U[0] * U[7]; U[8] = U[0] * U[4] - U[3] * U[1];
    V[2] = V[3] * V[7] - V[6] * V[4]; V[5] = V[6] * V[1] - V[0] * V[7]; V[8] = V[0] * V[4] - V[3] * V[1];
    double R[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) R[3 * r + c] = (U[3 * r] * V[3 * c] + U[3 * r + 1] * V[3 * c + 1]) + U[3 * r + 2] * V[3 * c + 2];
    for (int r = 0; r < 3; ++r) {
        const double t = dm[r] - ((R[3 * r] * sm[0] + R[3 * r + 1] * sm[1]) + R[3 * r + 2] * sm[2]);
        T[4 * r] = (float)R[3 * r]; T[4 * r + 1] = (float)R[3 * r + 1]; T[4 * r + 2] = (float)R[3 * r + 2]; T[4 * r + 3] = (float)t;
    }
    T[12] = 0.0f; T[13] = 0.0f; T[14] = 0.0f; T[15] = 1.0f;
}

// |T * src - tgt|^2 in float: Eigen's 4x4 * vec4 (columns scaled and added left to right), then squaredNorm
__host__ __device__ inline float transfer_sqd(const float T[16], float sx, float sy, float sz, float tx, float ty, float tz) {
    const float px = ((T[0] * sx + T[1] * sy) + T[2] * sz) + T[3];
    const float py = ((T[4] * sx + T[5] * sy) + T[6] * sz) + T[7];
    const float pz = ((T[8] * sx + T[9] * sy) + T[10] * sz) + T[11];
    const float dx = px - tx, dy = py - ty, dz = pz - tz;
    return (dx * dx + dy * dy) + dz * dz;
}

// one warp per hypothesis: model from its three pairs, inlier count over all correspondences
__global__ void __launch_bounds__(128)
ransac_score_kernel(const float4* __restrict__ src, const float4* __restrict__ tgt, unsigned n, const int* __restrict__ samples, unsigned n_hyp,
                    double thresh2, float* __restrict__ models, int* __restrict__ counts) {
    const unsigned lane = threadIdx.x & 31, h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (h >= n_hyp) return;
    float T[16];
    {
        double s[3][3], d[3][3];
        for (int i = 0; i < 3; ++i) {
            const int k = samples[3 * h + i];
            const float4 a = src[k], b = tgt[k];
            s[i][0] = a.x; s[i][1] = a.y; s[i][2] = a.z;
            d[i][0] = b.x; d[i][1] = b.y; d[i][2] = b.z;
        }
        umeyama3(s, d, T);  // every lane computes the same model (no divergence, no broadcast needed)
    }
    int cnt = 0;
    for (unsigned i = lane; i < n; i += 32) {
        const float4 a = src[i], b = tgt[i];
        if ((double)transfer_sqd(T, a.x, a.y, a.z, b.x, b.y, b.z) < thresh2) ++cnt;
    }
    cnt = warp_sum(cnt);
    if (lane == 0) counts[h] = cnt;
    if (lane < 16) models[16 * h + lane] = T[lane];
}

__global__ void ransac_select_kernel(const float4* __restrict__ src, const float4* __restrict__ tgt, unsigned n, const float* __restrict__ model,
                                     double thresh2, unsigned char* __restrict__ inlier) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float T[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) T[k] = model[k];
    const float4 a = src[i], b = tgt[i];
    inlier[i] = ((double)transfer_sqd(T, a.x, a.y, a.z, b.x, b.y, b.z) < thresh2) ? 1 : 0;
}

// ---- host: PCL's sample sequence ------------------------------------------------------------------------------------------
namespace {
struct Mt19937 {  // boost::mt19937 / std::mt19937 recurrence
    uint32_t s[624];
    int idx;
    explicit Mt19937(uint32_t seed) {
        s[0] = seed;
        for (int i = 1; i < 624; ++i) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    uint32_t next() {
        if (idx >= 624) {
            for (int i = 0; i < 624; ++i) {
                const uint32_t y = (s[i] & 0x80000000u) | (s[(i + 1) % 624] & 0x7FFFFFFFu);
                s[i] = s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
            }
            idx = 0;
        }
        uint32_t y = s[idx++];
        y ^= y >> 11; y ^= (y << 7) & 0x9D2C5680u; y ^= (y << 15) & 0xEFC60000u; y ^= y >> 18;
        return y;
    }
};

// pcl::eigen33 eigenvalues of a symmetric 3x3 in float are only needed for sample_dist_thresh; computed here in double
// by Jacobi and rounded (the threshold only steers which samples are retried)
void sym_eigenvalues(const double m_in[9], double w[3]) {
    double U[9], V[9];
    svd3_hestenes(m_in, U, w, V);  // symmetric positive semi-definite: singular values = eigenvalues
}
}  // namespace

// sample_dist_thresh of SampleConsensusModelRegistration::computeSampleDistanceThreshold over the correspondence sources
static double sample_dist_threshold(const float* src4, size_t n) {
    // computeMeanAndCovarianceMatrix: single pass, float accumulators, index order
    float a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t i = 0; i < n; ++i) {
        const float x = src4[4 * i], y = src4[4 * i + 1], z = src4[4 * i + 2];
        a[0] += x * x; a[1] += x * y; a[2] += x * z; a[3] += y * y; a[4] += y * z; a[5] += z * z; a[6] += x; a[7] += y; a[8] += z;
    }
    const float fn = (float)n;
    for (int k = 0; k < 9; ++k) a[k] /= fn;
    const double cov[9] = {(double)(a[0] - a[6] * a[6]), (double)(a[1] - a[6] * a[7]), (double)(a[2] - a[6] * a[8]),
                           (double)(a[1] - a[6] * a[7]), (double)(a[3] - a[7] * a[7]), (double)(a[4] - a[7] * a[8]),
                           (double)(a[2] - a[6] * a[8]), (double)(a[4] - a[7] * a[8]), (double)(a[5] - a[8] * a[8])};
    double w[3];
    sym_eigenvalues(cov, w);
    const float e0 = (float)w[0], e1 = (float)w[1], e2 = (float)w[2];
    double t = (double)((sqrtf(fmaxf(e0, 0.0f)) + sqrtf(fmaxf(e1, 0.0f))) + sqrtf(fmaxf(e2, 0.0f))) / 3.0;
    return t * t;
}

int ransac_run(Ctx* c, const float* src_xyz, const float* tgt_xyz, const int* pairs, size_t n_pairs, int max_iterations, double threshold,
               int* inlier_pairs_out, int* n_inliers_out, float* transform_out, int* iterations_out) {
    if (n_inliers_out) *n_inliers_out = 0;
    auto keep_all = [&]() {  // PCL: computeModel failed / fewer than 3 inliers -> all correspondences stay, identity
        if (inlier_pairs_out) memcpy(inlier_pairs_out, pairs, sizeof(int) * 2 * n_pairs);
        if (n_inliers_out) *n_inliers_out = (int)n_pairs;
        if (transform_out) for (int k = 0; k < 16; ++k) transform_out[k] = (k % 5 == 0) ? 1.0f : 0.0f;
        if (iterations_out) *iterations_out = 0;
        return BSHOT_OK;
    };
    if (n_pairs < 3) return keep_all();
    if (n_pairs > c->max_kp) { set_error("bshot_ransac: %zu correspondences > max_keypoints %zu", n_pairs, c->max_kp); return BSHOT_E_CAPACITY; }
    if (max_iterations < 1 || max_iterations > 100000) { set_error("bshot_ransac: bad max_iterations"); return BSHOT_E_INVALID; }
    const size_t n = n_pairs, n_hyp = (size_t)max_iterations + 1;
    // correspondence sources / targets in correspondence order (what indices_ / indices_tgt_ address)
    std::vector<float> s4(4 * n), t4(4 * n);
    for (size_t i = 0; i < n; ++i) {
        const int q = pairs[2 * i], m = pairs[2 * i + 1];
        for (int k = 0; k < 3; ++k) { s4[4 * i + k] = src_xyz[3 * (size_t)q + k]; t4[4 * i + k] = tgt_xyz[3 * (size_t)m + k]; }
        s4[4 * i + 3] = t4[4 * i + 3] = 1.0f;
    }
    // ---- PCL's sample sequence for iterations 0 .. max_iterations (independent of the scores) -------------------------------
    const double sdt = sample_dist_threshold(s4.data(), n);
    std::vector<int> shuffled(n), samples(3 * n_hyp);
    for (size_t i = 0; i < n; ++i) shuffled[i] = (int)i;
    Mt19937 rng(12345u);
    size_t n_valid = n_hyp;
    for (size_t h = 0; h < n_hyp; ++h) {
        bool good = false;
        for (int tries = 0; tries < 1000 && !good; ++tries) {  // max_sample_checks_
            for (size_t i = 0; i < 3; ++i) std::swap(shuffled[i], shuffled[i + ((rng.next() >> 1) % (n - i))]);
            auto d2 = [&](int a, int b) {
                const float dx = s4[4 * b] - s4[4 * a], dy = s4[4 * b + 1] - s4[4 * a + 1], dz = s4[4 * b + 2] - s4[4 * a + 2];
                return (double)(dx * dx + dy * dy + dz * dz);
            };
            good = d2(shuffled[0], shuffled[1]) > sdt && d2(shuffled[0], shuffled[2]) > sdt && d2(shuffled[1], shuffled[2]) > sdt;
        }
        if (!good) { n_valid = h; break; }  // "Could not select sample points": PCL stops the loop here
        samples[3 * h] = shuffled[0]; samples[3 * h + 1] = shuffled[1]; samples[3 * h + 2] = shuffled[2];
    }
    if (n_valid == 0) return keep_all();
    // ---- score every hypothesis on the device -----------------------------------------------------------------------------
    float4* d_src = reinterpret_cast<float4*>(c->d_gather);          // max_kp x 48 B scratch: sources | targets
    float4* d_tgt = d_src + n;
    int* d_samples = reinterpret_cast<int*>(c->d_partial);            // >= 2 x 256 x max_kp u64
    float* d_models = reinterpret_cast<float*>(d_samples + 3 * n_hyp);
    int* d_counts = reinterpret_cast<int*>(d_models + 16 * n_hyp);
    unsigned char* d_inlier = reinterpret_cast<unsigned char*>(d_counts + n_hyp);
    if ((3 + 16 + 1) * n_hyp * 4 + n > c->partial_cap * 8) { set_error("bshot_ransac: scratch too small for %d iterations", max_iterations); return BSHOT_E_CAPACITY; }
    BSHOT_CUDA_TRY(cudaMemcpyAsync(d_src, s4.data(), 16 * n, cudaMemcpyHostToDevice, c->stream));
    BSHOT_CUDA_TRY(cudaMemcpyAsync(d_tgt, t4.data(), 16 * n, cudaMemcpyHostToDevice, c->stream));
    BSHOT_CUDA_TRY(cudaMemcpyAsync(d_samples, samples.data(), sizeof(int) * 3 * n_valid, cudaMemcpyHostToDevice, c->stream));
    const double thresh2 = threshold * threshold;
    ransac_score_kernel<<<(unsigned)((n_valid * 32 + 127) / 128), 128, 0, c->stream>>>(d_src, d_tgt, (unsigned)n, d_samples, (unsigned)n_valid, thresh2, d_models,
                                                                                     d_counts);
    count_launch(c);
    BSHOT_TRY(check_launch("ransac_score_kernel"));
    std::vector<int> counts(n_valid);
    BSHOT_CUDA_TRY(cudaMemcpyAsync(counts.data(), d_counts, sizeof(int) * n_valid, cudaMemcpyDeviceToHost, c->stream));
    BSHOT_CUDA_TRY(cudaStreamSynchronize(c->stream));
    // ---- RandomSampleConsensus::computeModel replayed over the scores --------------------------------------------------------
    int iterations = 0, best = -1, n_best = -2147483647;
    double k = 1.0;
    const double log_probability = log(1.0 - 0.99), one_over_indices = 1.0 / (double)n;
    while ((double)iterations < k) {
        if ((size_t)iterations >= n_valid) break;  // no sample could be drawn: PCL breaks out
        const int cnt = counts[iterations];
        if (cnt > n_best) {
            n_best = cnt;
            best = iterations;
            const double w = (double)n_best * one_over_indices;
            double p_no_outliers = 1.0 - pow(w, 3.0);
            p_no_outliers = std::max(2.220446049250313e-16, p_no_outliers);
            p_no_outliers = std::min(1.0 - 2.220446049250313e-16, p_no_outliers);
            k = log_probability / log(p_no_outliers);
        }
        ++iterations;
        if (iterations > max_iterations) break;
    }
    if (best < 0) return keep_all();
    // ---- inliers of the best model, original order ---------------------------------------------------------------------------
    ransac_select_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_src, d_tgt, (unsigned)n, d_models + 16 * (size_t)best, thresh2, d_inlier);
    count_launch(c);
    BSHOT_TRY(check_launch("ransac_select_kernel"));
    std::vector<unsigned char> inl(n);
    float model[16];
    BSHOT_CUDA_TRY(cudaMemcpyAsync(inl.data(), d_inlier, n, cudaMemcpyDeviceToHost, c->stream));
    BSHOT_CUDA_TRY(cudaMemcpyAsync(model, d_models + 16 * (size_t)best, sizeof(model), cudaMemcpyDeviceToHost, c->stream));
    BSHOT_CUDA_TRY(cudaStreamSynchronize(c->stream));
    size_t m = 0;
    for (size_t i = 0; i < n; ++i) m += inl[i];
    if (m < 3) return keep_all();
    m = 0;
    for (size_t i = 0; i < n; ++i)
        if (inl[i]) {
            if (inlier_pairs_out) { inlier_pairs_out[2 * m] = pairs[2 * i]; inlier_pairs_out[2 * m + 1] = pairs[2 * i + 1]; }
            ++m;
        }
    if (n_inliers_out) *n_inliers_out = (int)m;
    if (transform_out) memcpy(transform_out, model, sizeof(model));
    if (iterations_out) *iterations_out = iterations;
    return BSHOT_OK;
}

}  // namespace bshot

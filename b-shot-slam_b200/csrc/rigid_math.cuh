// rigid_math.cuh -- small dense pieces shared by the RANSAC and ICP kernels (both compiled with -fmad=false so that the
// CPU restatement in the test oracle reproduces them bit for bit).
#pragma once
#include <math.h>

namespace bshot {

// ---- 3x3 SVD by one-sided Jacobi (Hestenes), double; shared by host and device ---------------------------------------
// A = U diag(s) V^T with s[0] >= s[1] >= s[2] >= 0.  Columns of U for (near) zero singular values are completed to a
// right-handed basis.  Only +, -, *, /, sqrt: reproducible bit for bit without FMA contraction.
__host__ __device__ inline void svd3_hestenes(const double a_in[9], double U[9], double s[3], double V[9]) {
    double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) a[r][c] = a_in[3 * r + c];
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int r = 0; r < 3; ++r) { alpha += a[r][p] * a[r][p]; beta += a[r][q] * a[r][q]; gamma += a[r][p] * a[r][q]; }
                if (gamma == 0.0) continue;
                const double lim = 1e-30 * (alpha * beta);
                if (gamma * gamma <= lim) continue;
                off += gamma * gamma;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                for (int r = 0; r < 3; ++r) {
                    const double x = a[r][p], y = a[r][q];
                    a[r][p] = c * x - sn * y;
                    a[r][q] = sn * x + c * y;
                    const double vx = v[r][p], vy = v[r][q];
                    v[r][p] = c * vx - sn * vy;
                    v[r][q] = sn * vx + c * vy;
                }
            }
        if (off == 0.0) break;
    }
    double n[3];
    int o[3] = {0, 1, 2};
    for (int c = 0; c < 3; ++c) n[c] = sqrt(a[0][c] * a[0][c] + a[1][c] * a[1][c] + a[2][c] * a[2][c]);
    if (n[o[0]] < n[o[1]]) { int t = o[0]; o[0] = o[1]; o[1] = t; }
    if (n[o[1]] < n[o[2]]) { int t = o[1]; o[1] = o[2]; o[2] = t; }
    if (n[o[0]] < n[o[1]]) { int t = o[0]; o[0] = o[1]; o[1] = t; }
    for (int k = 0; k < 3; ++k) {
        s[k] = n[o[k]];
        for (int r = 0; r < 3; ++r) { V[3 * r + k] = v[r][o[k]]; U[3 * r + k] = (s[k] > 0.0) ? a[r][o[k]] / s[k] : 0.0; }
    }
}

}  // namespace bshot

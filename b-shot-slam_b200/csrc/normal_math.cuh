// normal_math.cuh -- pcl::computePointNormal arithmetic shared by the normal kernels:
// covariance from the nine single-pass sums, smallest eigenpair by the closed-form pcl::eigen33, curvature,
// flip towards the origin (include/bshot_bits.h:77,83).  fp32, products evaluated left to right, no FMA contraction.
#pragma once
#include "common.cuh"

namespace bshot {

__device__ __forceinline__ void compute_roots2(float b, float c, float roots[3]) {
    roots[0] = 0.0f;
    float d = (float)((double)__fmul_rn(b, b) - 4.0 * (double)c);
    if (d < 0.0f) d = 0.0f;
    const float sd = sqrtf(d);
    roots[2] = __fmul_rn(0.5f, __fadd_rn(b, sd));
    roots[1] = __fmul_rn(0.5f, __fsub_rn(b, sd));
}

// pcl::computeRoots (fp32, trigonometric closed form); products evaluated left to right, no FMA
__device__ inline void compute_roots(const float m[9], float roots[3]) {
    const float m00 = m[0], m01 = m[1], m02 = m[2], m11 = m[4], m12 = m[5], m22 = m[8];
    float c0 = __fmul_rn(__fmul_rn(m00, m11), m22);
    c0 = __fadd_rn(c0, __fmul_rn(__fmul_rn(__fmul_rn(2.0f, m01), m02), m12));
    c0 = __fsub_rn(c0, __fmul_rn(__fmul_rn(m00, m12), m12));
    c0 = __fsub_rn(c0, __fmul_rn(__fmul_rn(m11, m02), m02));
    c0 = __fsub_rn(c0, __fmul_rn(__fmul_rn(m22, m01), m01));
    float c1 = __fsub_rn(__fmul_rn(m00, m11), __fmul_rn(m01, m01));
    c1 = __fadd_rn(c1, __fmul_rn(m00, m22));
    c1 = __fsub_rn(c1, __fmul_rn(m02, m02));
    c1 = __fadd_rn(c1, __fmul_rn(m11, m22));
    c1 = __fsub_rn(c1, __fmul_rn(m12, m12));
    const float c2 = __fadd_rn(__fadd_rn(m00, m11), m22);
    if (fabsf(c0) < 1.1920929e-07f) {
        compute_roots2(c2, c1, roots);
        return;
    }
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = sqrtf(3.0f);
    const float c2_over_3 = __fmul_rn(c2, s_inv3);
    float a_over_3 = __fmul_rn(__fsub_rn(c1, __fmul_rn(c2, c2_over_3)), s_inv3);
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    const float inner = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, c2_over_3), c2_over_3), c1);
    const float half_b = __fmul_rn(0.5f, __fadd_rn(c0, __fmul_rn(c2_over_3, inner)));
    float qv = __fadd_rn(__fmul_rn(half_b, half_b), __fmul_rn(__fmul_rn(a_over_3, a_over_3), a_over_3));
    if (qv > 0.0f) qv = 0.0f;
    const float rho = sqrtf(-a_over_3);
    const float theta = __fmul_rn(atan2f(sqrtf(-qv), half_b), s_inv3);
    const float cos_theta = cosf(theta);
    const float sin_theta = sinf(theta);
    roots[0] = __fadd_rn(c2_over_3, __fmul_rn(__fmul_rn(2.0f, rho), cos_theta));
    roots[1] = __fsub_rn(c2_over_3, __fmul_rn(rho, __fadd_rn(cos_theta, __fmul_rn(s_sqrt3, sin_theta))));
    roots[2] = __fsub_rn(c2_over_3, __fmul_rn(rho, __fsub_rn(cos_theta, __fmul_rn(s_sqrt3, sin_theta))));
    if (roots[0] >= roots[1]) { const float t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
    if (roots[1] >= roots[2]) {
        const float t = roots[1]; roots[1] = roots[2]; roots[2] = t;
        if (roots[0] >= roots[1]) { const float u = roots[0]; roots[0] = roots[1]; roots[1] = u; }
    }
    if (roots[0] <= 0.0f) compute_roots2(c2, c1, roots);
}

__device__ __forceinline__ void cross_rn(const float* a, const float* b, float* o) {
    o[0] = __fsub_rn(__fmul_rn(a[1], b[2]), __fmul_rn(a[2], b[1]));
    o[1] = __fsub_rn(__fmul_rn(a[2], b[0]), __fmul_rn(a[0], b[2]));
    o[2] = __fsub_rn(__fmul_rn(a[0], b[1]), __fmul_rn(a[1], b[0]));
}

// pcl::eigen33(mat, eigenvalue, eigenvector): smallest eigenpair
__device__ inline void eigen33_smallest(const float mat[9], float& eigenvalue, float evec[3]) {
    float scale = 0.0f;
#pragma unroll
    for (int k = 0; k < 9; ++k) scale = fmaxf(scale, fabsf(mat[k]));
    if (scale <= 1.17549435e-38f) scale = 1.0f;
    float s[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) s[k] = mat[k] / scale;
    float roots[3];
    compute_roots(s, roots);
    eigenvalue = __fmul_rn(roots[0], scale);
    s[0] = __fsub_rn(s[0], roots[0]); s[4] = __fsub_rn(s[4], roots[0]); s[8] = __fsub_rn(s[8], roots[0]);
    float v1[3], v2[3], v3[3];
    cross_rn(s, s + 3, v1);
    cross_rn(s, s + 6, v2);
    cross_rn(s + 3, s + 6, v3);
    const float l1 = dot3_rn(v1[0], v1[1], v1[2], v1[0], v1[1], v1[2]);
    const float l2 = dot3_rn(v2[0], v2[1], v2[2], v2[0], v2[1], v2[2]);
    const float l3 = dot3_rn(v3[0], v3[1], v3[2], v3[0], v3[1], v3[2]);
    const float* v; float l;
    if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; }
    else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
    else { v = v3; l = l3; }
    const float inv = sqrtf(l);
    evec[0] = v[0] / inv; evec[1] = v[1] / inv; evec[2] = v[2] / inv;
}

// a[0..8] = sums of x*x, x*y, x*z, y*y, y*z, z*z, x, y, z over the n selected neighbours (fp32, neighbour order);
// q = the query point.  Returns (nx, ny, nz, curvature); NaN when n < 3.
__device__ __forceinline__ float4 normal_from_sums(const float s[9], int n, float qx, float qy, float qz) {
    const float nanf_ = __int_as_float(0x7FC00000);
    float4 o = make_float4(nanf_, nanf_, nanf_, nanf_);
    if (n < 3) return o;  // n == 0: include/bshot_bits.h:67-74 ; n < 3: computePointNormal guard
    const float fn = (float)n;
    float a[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) a[k] = s[k] / fn;
    float cov[9];
    cov[0] = __fsub_rn(a[0], __fmul_rn(a[6], a[6]));
    cov[1] = __fsub_rn(a[1], __fmul_rn(a[6], a[7]));
    cov[2] = __fsub_rn(a[2], __fmul_rn(a[6], a[8]));
    cov[4] = __fsub_rn(a[3], __fmul_rn(a[7], a[7]));
    cov[5] = __fsub_rn(a[4], __fmul_rn(a[7], a[8]));
    cov[8] = __fsub_rn(a[5], __fmul_rn(a[8], a[8]));
    cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
    float ev, nv[3];
    eigen33_smallest(cov, ev, nv);
    const float eig_sum = __fadd_rn(__fadd_rn(cov[0], cov[4]), cov[8]);
    o.w = (eig_sum != 0.0f) ? fabsf(ev / eig_sum) : 0.0f;
    const float vx = 0.0f - qx, vy = 0.0f - qy, vz = 0.0f - qz;
    const float cos_theta = __fadd_rn(__fadd_rn(__fmul_rn(vx, nv[0]), __fmul_rn(vy, nv[1])), __fmul_rn(vz, nv[2]));
    if (cos_theta < 0.0f) { nv[0] = -nv[0]; nv[1] = -nv[1]; nv[2] = -nv[2]; }
    o.x = nv[0]; o.y = nv[1]; o.z = nv[2];
    return o;
}

}  // namespace bshot

// Device-side SO(3) / small-matrix helpers for the sm_100a kernels.
// Small-angle branches follow the reference exactly (include/utilities/geometry.h:30-37,131-166):
//   so3_rightJacobian -> identity below 1e-5, exp_so3 -> first order below 1e-9, log_so3 -> first order below 1e-9.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define SDV_DEV __device__ __forceinline__

namespace sdv {

SDV_DEV void mat3_mul(const double *A, const double *B, double *C) { // C = A*B (3x3 row-major)
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
SDV_DEV void mat3_mulT(const double *A, const double *B, double *C) { // C = A*B^T
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3] * B[j * 3] + A[i * 3 + 1] * B[j * 3 + 1] + A[i * 3 + 2] * B[j * 3 + 2];
}
SDV_DEV void matT3_mul(const double *A, const double *B, double *C) { // C = A^T*B
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
SDV_DEV void mat3_vec(const double *A, const double *v, double *o) {
#pragma unroll
    for (int i = 0; i < 3; i++) o[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}
SDV_DEV void matT3_vec(const double *A, const double *v, double *o) {
#pragma unroll
    for (int i = 0; i < 3; i++) o[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
}
SDV_DEV void skew3(const double *w, double *S) {
    S[0] = 0;     S[1] = -w[2]; S[2] = w[1];
    S[3] = w[2];  S[4] = 0;     S[5] = -w[0];
    S[6] = -w[1]; S[7] = w[0];  S[8] = 0;
}
SDV_DEV void cross3(const double *a, const double *b, double *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
SDV_DEV double norm3(const double *a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

SDV_DEV void exp_so3(const double *v, double *R) { // geometry.h:131-147
    double angle = norm3(v);
    if (angle < 1e-9) {
        skew3(v, R);
        R[0] += 1.0; R[4] += 1.0; R[8] += 1.0;
        return;
    }
    double ax[3] = {v[0] / angle, v[1] / angle, v[2] / angle};
    double K[9], K2[9];
    skew3(ax, K);
    mat3_mul(K, K, K2);
    double s, c;
    sincos(angle, &s, &c);
#pragma unroll
    for (int i = 0; i < 9; i++) R[i] = (1.0 - c) * K2[i] + s * K[i];
    R[0] += 1.0; R[4] += 1.0; R[8] += 1.0;
}

SDV_DEV void log_so3(const double *M, double *phi) { // geometry.h:149-166
    double cos_angle = 0.5 * (M[0] + M[4] + M[8]) - 0.5;
    cos_angle = fmin(fmax(cos_angle, -1.0), 1.0);
    double angle = acos(cos_angle);
    double w[3] = {M[7] - M[5], M[2] - M[6], M[3] - M[1]};
    double sa = sin(angle);
    double f = (fabs(sa) < 1e-9 || angle < 1e-9) ? 0.5 : 0.5 * angle / sa;
    phi[0] = f * w[0]; phi[1] = f * w[1]; phi[2] = f * w[2];
}

SDV_DEV void right_jacobian(const double *w, double *J) { // geometry.h:30-37
    double n = norm3(w);
    if (n < 1e-5) {
#pragma unroll
        for (int i = 0; i < 9; i++) J[i] = 0.0;
        J[0] = J[4] = J[8] = 1.0;
        return;
    }
    double S[9], S2[9];
    skew3(w, S);
    mat3_mul(S, S, S2);
    double s, c;
    sincos(n, &s, &c);
    double a = (1.0 - c) / (n * n), b = (n - s) / (n * n * n);
#pragma unroll
    for (int i = 0; i < 9; i++) J[i] = -a * S[i] + b * S2[i];
    J[0] += 1.0; J[4] += 1.0; J[8] += 1.0;
}

SDV_DEV void inverse3(const double *m, double *r) { // cofactor inverse
    double c00 = m[4] * m[8] - m[5] * m[7];
    double c01 = m[5] * m[6] - m[3] * m[8];
    double c02 = m[3] * m[7] - m[4] * m[6];
    double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    double id = 1.0 / det;
    r[0] = c00 * id;
    r[3] = c01 * id;
    r[6] = c02 * id;
    r[1] = (m[2] * m[7] - m[1] * m[8]) * id;
    r[4] = (m[0] * m[8] - m[2] * m[6]) * id;
    r[7] = (m[1] * m[6] - m[0] * m[7]) * id;
    r[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    r[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    r[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

// Symmetric 3x3 inverse from the 6 unique entries (xx,xy,xz,yy,yz,zz); returns false when not positive definite
// (checked through the leading minors, which is what a 3x3 Cholesky would detect).
SDV_DEV bool sym3_inverse(const double *h, double *inv) {
    double a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5];
    double m1 = a, m2 = a * d - b * b;
    double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
    double det = a * c00 + b * c01 + c * c02;
    if (!(m1 > 0.0) || !(m2 > 0.0) || !(det > 0.0)) return false;
    double id = 1.0 / det;
    inv[0] = c00 * id;
    inv[1] = c01 * id;
    inv[2] = c02 * id;
    inv[3] = (a * f - c * c) * id;
    inv[4] = (b * c - a * e) * id;
    inv[5] = (a * d - b * b) * id;
    return true;
}

SDV_DEV double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// atomic max for non-negative doubles (bit pattern order == numeric order)
SDV_DEV void atomic_max_nonneg(double *addr, double v) {
    atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

} // namespace sdv

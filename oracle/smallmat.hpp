// TEST INFRASTRUCTURE — part of the CPU oracle (see oracle/README.md). Not shipped, never on the product path.
//
// Minimal fixed-size row-major matrix type so that the oracle can restate the reference's Eigen
// expressions one-to-one without Eigen (which is not installed in this image).
#pragma once
#include <cmath>
#include <cstring>

namespace orc {

template <int R, int C> struct Mat {
    double d[R * C];
    double &operator()(int i, int j) { return d[i * C + j]; }
    double operator()(int i, int j) const { return d[i * C + j]; }
    double &operator[](int i) { return d[i]; }
    double operator[](int i) const { return d[i]; }
    static Mat Zero() {
        Mat m;
        for (int i = 0; i < R * C; i++) m.d[i] = 0.0;
        return m;
    }
    static Mat Identity() {
        Mat m = Zero();
        for (int i = 0; i < (R < C ? R : C); i++) m(i, i) = 1.0;
        return m;
    }
    static Mat From(const double *p) {
        Mat m;
        std::memcpy(m.d, p, sizeof(double) * R * C);
        return m;
    }
    void to(double *p) const { std::memcpy(p, d, sizeof(double) * R * C); }
    Mat<C, R> T() const {
        Mat<C, R> t;
        for (int i = 0; i < R; i++)
            for (int j = 0; j < C; j++) t(j, i) = (*this)(i, j);
        return t;
    }
    double norm() const {
        double s = 0;
        for (int i = 0; i < R * C; i++) s += d[i] * d[i];
        return std::sqrt(s);
    }
    double dot(const Mat &o) const {
        double s = 0;
        for (int i = 0; i < R * C; i++) s += d[i] * o.d[i];
        return s;
    }
    double trace() const {
        double s = 0;
        for (int i = 0; i < (R < C ? R : C); i++) s += (*this)(i, i);
        return s;
    }
    template <int BR, int BC> Mat<BR, BC> block(int r0, int c0) const {
        Mat<BR, BC> b;
        for (int i = 0; i < BR; i++)
            for (int j = 0; j < BC; j++) b(i, j) = (*this)(r0 + i, c0 + j);
        return b;
    }
    template <int BR, int BC> void setBlock(int r0, int c0, const Mat<BR, BC> &b) {
        for (int i = 0; i < BR; i++)
            for (int j = 0; j < BC; j++) (*this)(r0 + i, c0 + j) = b(i, j);
    }
};

template <int R, int K, int C> inline Mat<R, C> operator*(const Mat<R, K> &a, const Mat<K, C> &b) {
    Mat<R, C> m;
    for (int i = 0; i < R; i++)
        for (int j = 0; j < C; j++) {
            double s = 0;
            for (int k = 0; k < K; k++) s += a(i, k) * b(k, j);
            m(i, j) = s;
        }
    return m;
}
template <int R, int C> inline Mat<R, C> operator+(const Mat<R, C> &a, const Mat<R, C> &b) {
    Mat<R, C> m;
    for (int i = 0; i < R * C; i++) m.d[i] = a.d[i] + b.d[i];
    return m;
}
template <int R, int C> inline Mat<R, C> operator-(const Mat<R, C> &a, const Mat<R, C> &b) {
    Mat<R, C> m;
    for (int i = 0; i < R * C; i++) m.d[i] = a.d[i] - b.d[i];
    return m;
}
template <int R, int C> inline Mat<R, C> operator-(const Mat<R, C> &a) {
    Mat<R, C> m;
    for (int i = 0; i < R * C; i++) m.d[i] = -a.d[i];
    return m;
}
template <int R, int C> inline Mat<R, C> operator*(double s, const Mat<R, C> &a) {
    Mat<R, C> m;
    for (int i = 0; i < R * C; i++) m.d[i] = s * a.d[i];
    return m;
}
template <int R, int C> inline Mat<R, C> operator*(const Mat<R, C> &a, double s) { return s * a; }
template <int R, int C> inline Mat<R, C> operator/(const Mat<R, C> &a, double s) {
    Mat<R, C> m;
    for (int i = 0; i < R * C; i++) m.d[i] = a.d[i] / s;
    return m;
}

using M3 = Mat<3, 3>;
using V3 = Mat<3, 1>;

inline V3 vec3(double x, double y, double z) {
    V3 v;
    v[0] = x;
    v[1] = y;
    v[2] = z;
    return v;
}
inline V3 cross(const V3 &a, const V3 &b) {
    return vec3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
inline M3 inverse3(const M3 &m) {
    // cofactor inverse (what Eigen uses for fixed 3x3)
    M3 r;
    double c00 = m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1);
    double c01 = m(1, 2) * m(2, 0) - m(1, 0) * m(2, 2);
    double c02 = m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0);
    double det = m(0, 0) * c00 + m(0, 1) * c01 + m(0, 2) * c02;
    double id = 1.0 / det;
    r(0, 0) = c00 * id;
    r(1, 0) = c01 * id;
    r(2, 0) = c02 * id;
    r(0, 1) = (m(0, 2) * m(2, 1) - m(0, 1) * m(2, 2)) * id;
    r(1, 1) = (m(0, 0) * m(2, 2) - m(0, 2) * m(2, 0)) * id;
    r(2, 1) = (m(0, 1) * m(2, 0) - m(0, 0) * m(2, 1)) * id;
    r(0, 2) = (m(0, 1) * m(1, 2) - m(0, 2) * m(1, 1)) * id;
    r(1, 2) = (m(0, 2) * m(1, 0) - m(0, 0) * m(1, 2)) * id;
    r(2, 2) = (m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0)) * id;
    return r;
}

// General N x N inverse by LU with partial pivoting (Eigen's MatrixXd::inverse() path).
template <int N> inline bool inverseLU(const Mat<N, N> &a, Mat<N, N> &out) {
    double m[N][2 * N];
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
            m[i][j] = a(i, j);
            m[i][N + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < N; c++) {
        int p = c;
        double best = std::fabs(m[c][c]);
        for (int r = c + 1; r < N; r++)
            if (std::fabs(m[r][c]) > best) {
                best = std::fabs(m[r][c]);
                p = r;
            }
        if (best == 0.0) return false;
        if (p != c)
            for (int j = 0; j < 2 * N; j++) {
                double t = m[c][j];
                m[c][j] = m[p][j];
                m[p][j] = t;
            }
        double inv = 1.0 / m[c][c];
        for (int r = c + 1; r < N; r++) {
            double f = m[r][c] * inv;
            if (f == 0.0) continue;
            for (int j = c; j < 2 * N; j++) m[r][j] -= f * m[c][j];
        }
    }
    for (int c = N - 1; c >= 0; c--) {
        double inv = 1.0 / m[c][c];
        for (int j = 0; j < 2 * N; j++) m[c][j] *= inv;
        for (int r = 0; r < c; r++) {
            double f = m[r][c];
            if (f == 0.0) continue;
            for (int j = 0; j < 2 * N; j++) m[r][j] -= f * m[c][j];
        }
    }
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) out(i, j) = m[i][N + j];
    return true;
}

// Lower Cholesky factor L (A = L L^T); returns false if not positive definite.
template <int N> inline bool choleskyL(const Mat<N, N> &a, Mat<N, N> &L) {
    L = Mat<N, N>::Zero();
    for (int j = 0; j < N; j++) {
        double s = a(j, j);
        for (int k = 0; k < j; k++) s -= L(j, k) * L(j, k);
        if (!(s > 0.0)) return false;
        double d = std::sqrt(s);
        L(j, j) = d;
        for (int i = j + 1; i < N; i++) {
            double t = a(i, j);
            for (int k = 0; k < j; k++) t -= L(i, k) * L(j, k);
            L(i, j) = t / d;
        }
    }
    return true;
}

// Rigid transform [R | t], the top 3 rows of Eigen::Affine3d.
struct Aff {
    M3 R;
    V3 t;
    static Aff Identity() { return {M3::Identity(), V3::Zero()}; }
    static Aff From(const double *p) { // 3x4 row-major
        Aff a;
        for (int i = 0; i < 3; i++) {
            for (int j = 0; j < 3; j++) a.R(i, j) = p[i * 4 + j];
            a.t[i] = p[i * 4 + 3];
        }
        return a;
    }
    void to(double *p) const {
        for (int i = 0; i < 3; i++) {
            for (int j = 0; j < 3; j++) p[i * 4 + j] = R(i, j);
            p[i * 4 + 3] = t[i];
        }
    }
    Aff operator*(const Aff &o) const { return {R * o.R, R * o.t + t}; }
    V3 operator*(const V3 &p) const { return R * p + t; }
    Aff inverse() const {
        M3 Rt = R.T();
        return {Rt, -(Rt * t)};
    }
};

} // namespace orc

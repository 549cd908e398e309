// TEST INFRASTRUCTURE (see oracle/README.md).  CPU restatement of the dense core of the marginal-prior construction of the
// reference — the checker of sdv_marginalize / sdv_schur_prior.
//
//   Marginalization::computeInformationAndGradient   cpp/src/optimizers/marginalization.cpp:145-211
//   Marginalization::computeSchurComplement          cpp/src/optimizers/marginalization.cpp:213-265
//   Marginalization::rankReveallingDecomposition     cpp/src/optimizers/marginalization.cpp:318-342
//   Marginalization::computeJacobiansAndResiduals    cpp/src/optimizers/marginalization.cpp:516-530
//
// Eigen::SelfAdjointEigenSolver (not in the image) is replaced by a cyclic Jacobi eigen-solver: eigenvalues agree to
// rounding, eigenvectors only up to sign / rotation inside degenerate eigenspaces — so J_m and r0 are compared through the
// invariants J_m^T J_m = A_k (on its range) and J_m^T r0 = -b_k, never entry by entry.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace orc {

// Symmetric eigen-decomposition A = V diag(w) V^T by Householder tridiagonalisation followed by the implicit QL iteration
// (the classical tred2 / tql2 pair of the Handbook for Automatic Computation — the same two stages Eigen's
// SelfAdjointEigenSolver runs, marginalization.cpp:229,323); row-major n x n, eigenvalues ascending like Eigen.
// O(n^3) with a small constant: the chained priors of a sliding window reach n ~ 400, where the cyclic Jacobi below needs
// tens of seconds.  jacobi_eig stays as the independent cross-check (tests/test_oracle_marginalization.py).
inline void sym_eig(int n, const double *A_in, std::vector<double> &w, std::vector<double> &Vout) {
    std::vector<double> V(A_in, A_in + (size_t)n * n), d(n, 0.0), e(n, 0.0);
    auto at = [&](int i, int j) -> double & { return V[(size_t)i * n + j]; };
    if (n == 0) {
        w.clear();
        Vout.clear();
        return;
    }
    for (int j = 0; j < n; j++) d[j] = at(n - 1, j);
    for (int i = n - 1; i > 0; i--) { // Householder reduction to tridiagonal form
        double scale = 0.0, h = 0.0;
        for (int k = 0; k < i; k++) scale += std::fabs(d[k]);
        if (scale == 0.0) {
            e[i] = d[i - 1];
            for (int j = 0; j < i; j++) {
                d[j] = at(i - 1, j);
                at(i, j) = 0.0;
                at(j, i) = 0.0;
            }
        } else {
            for (int k = 0; k < i; k++) {
                d[k] /= scale;
                h += d[k] * d[k];
            }
            double f = d[i - 1];
            double g = std::sqrt(h);
            if (f > 0) g = -g;
            e[i] = scale * g;
            h -= f * g;
            d[i - 1] = f - g;
            for (int j = 0; j < i; j++) e[j] = 0.0;
            for (int j = 0; j < i; j++) { // similarity transformation of the remaining columns
                f = d[j];
                at(j, i) = f;
                g = e[j] + at(j, j) * f;
                for (int k = j + 1; k <= i - 1; k++) {
                    g += at(k, j) * d[k];
                    e[k] += at(k, j) * f;
                }
                e[j] = g;
            }
            f = 0.0;
            for (int j = 0; j < i; j++) {
                e[j] /= h;
                f += e[j] * d[j];
            }
            const double hh = f / (h + h);
            for (int j = 0; j < i; j++) e[j] -= hh * d[j];
            for (int j = 0; j < i; j++) {
                f = d[j];
                g = e[j];
                for (int k = j; k <= i - 1; k++) at(k, j) -= f * e[k] + g * d[k];
                d[j] = at(i - 1, j);
                at(i, j) = 0.0;
            }
        }
        d[i] = h;
    }
    for (int i = 0; i < n - 1; i++) { // accumulate the transformations
        at(n - 1, i) = at(i, i);
        at(i, i) = 1.0;
        const double h = d[i + 1];
        if (h != 0.0) {
            for (int k = 0; k <= i; k++) d[k] = at(k, i + 1) / h;
            for (int j = 0; j <= i; j++) {
                double g = 0.0;
                for (int k = 0; k <= i; k++) g += at(k, i + 1) * at(k, j);
                for (int k = 0; k <= i; k++) at(k, j) -= g * d[k];
            }
        }
        for (int k = 0; k <= i; k++) at(k, i + 1) = 0.0;
    }
    for (int j = 0; j < n; j++) {
        d[j] = at(n - 1, j);
        at(n - 1, j) = 0.0;
    }
    at(n - 1, n - 1) = 1.0;
    e[0] = 0.0;
    // implicit QL on the tridiagonal matrix (d, e)
    for (int i = 1; i < n; i++) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    double f = 0.0, tst1 = 0.0;
    const double eps = std::ldexp(1.0, -52);
    for (int l = 0; l < n; l++) {
        tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
        int m = l;
        while (m < n - 1 && std::fabs(e[m]) > eps * tst1) m++;
        if (m > l) {
            int iter = 0;
            do {
                if (++iter > 200) break;
                double g = d[l];
                double p = (d[l + 1] - g) / (2.0 * e[l]);
                double r = std::hypot(p, 1.0);
                if (p < 0) r = -r;
                d[l] = e[l] / (p + r);
                d[l + 1] = e[l] * (p + r);
                const double dl1 = d[l + 1];
                double h = g - d[l];
                for (int i = l + 2; i < n; i++) d[i] -= h;
                f += h;
                p = d[m];
                double c = 1.0, c2 = c, c3 = c, s = 0.0, s2 = 0.0;
                const double el1 = e[l + 1];
                for (int i = m - 1; i >= l; i--) {
                    c3 = c2;
                    c2 = c;
                    s2 = s;
                    g = c * e[i];
                    h = c * p;
                    r = std::hypot(p, e[i]);
                    e[i + 1] = s * r;
                    s = e[i] / r;
                    c = p / r;
                    p = c * d[i] - s * g;
                    d[i + 1] = h + s * (c * g + s * d[i]);
                    for (int k = 0; k < n; k++) {
                        h = at(k, i + 1);
                        at(k, i + 1) = s * at(k, i) + c * h;
                        at(k, i) = c * at(k, i) - s * h;
                    }
                }
                p = -s * s2 * c3 * el1 * e[l] / dl1;
                e[l] = s * p;
                d[l] = c * p;
            } while (std::fabs(e[l]) > eps * tst1);
        }
        d[l] += f;
        e[l] = 0.0;
    }
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return d[a] < d[b]; });
    w.resize(n);
    Vout.assign((size_t)n * n, 0.0);
    for (int c = 0; c < n; c++) {
        w[c] = d[order[c]];
        for (int r = 0; r < n; r++) Vout[(size_t)r * n + c] = at(r, order[c]);
    }
}

// Symmetric eigen-decomposition A = V diag(w) V^T (cyclic Jacobi, row-major n x n, eigenvalues ascending like Eigen).
inline void jacobi_eig(int n, const double *A_in, std::vector<double> &w, std::vector<double> &V) {
    std::vector<double> A(A_in, A_in + (size_t)n * n);
    V.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) V[(size_t)i * n + i] = 1.0;
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) (i == j ? diag : off) += A[(size_t)i * n + j] * A[(size_t)i * n + j];
        if (off <= 1e-30 * (diag + 1e-300)) break;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = A[(size_t)p * n + q];
                if (apq == 0.0) continue;
                const double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) { // A <- A J
                    const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
                    A[(size_t)k * n + p] = c * akp - s * akq;
                    A[(size_t)k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) { // A <- J^T A
                    const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
                    A[(size_t)p * n + k] = c * apk - s * aqk;
                    A[(size_t)q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
                    V[(size_t)k * n + p] = c * vkp - s * vkq;
                    V[(size_t)k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return A[(size_t)a * n + a] < A[(size_t)b * n + b]; });
    w.resize(n);
    std::vector<double> Vs((size_t)n * n);
    for (int c = 0; c < n; c++) {
        w[c] = A[(size_t)order[c] * n + order[c]];
        for (int r = 0; r < n; r++) Vs[(size_t)r * n + c] = V[(size_t)r * n + order[c]];
    }
    V.swap(Vs);
}

// marginalization.cpp:213-265 + 318-342 + 516-530.  A is (m+n) x (m+n) row-major with the m marginalised parameters first,
// b has m+n entries.  Returns false when n < 4 (:215).  Outputs: Ak [n*n], bk [n], n_full, U [n * n_full] (column k = k-th
// kept eigenvector), Lambda [n_full], Jm [n_full * n] = Lambda^1/2 U^T, r0 [n_full] = -Lambda^-1/2 U^T bk.
inline bool schur_prior(int m, int n, const double *A, const double *b, double eps, std::vector<double> &Ak, std::vector<double> &bk,
                        int &n_full, std::vector<double> &U, std::vector<double> &Lambda, std::vector<double> &Jm, std::vector<double> &r0) {
    if (n < 4) return false;
    const int N = m + n;
    // Amm = 0.5 (A_mm + A_mm^T); pseudo-inverse through the eigen-decomposition, eigenvalues <= eps dropped (:234-240)
    std::vector<double> Amm((size_t)m * m), wm, Vm;
    for (int i = 0; i < m; i++)
        for (int j = 0; j < m; j++) Amm[(size_t)i * m + j] = 0.5 * (A[(size_t)i * N + j] + A[(size_t)j * N + i]);
    sym_eig(m, Amm.data(), wm, Vm);
    // Ak = Arr - Arm Amm^+ Arm^T, bk = brr - Arm Amm^+ bmm   (:242-248); Arm = A.block(m, 0, n, m).  With
    // Amm^+ = V diag(1/w) V^T the products are formed as (Arm V) diag(1/w) (Arm V)^T: the same algebra, symmetric by
    // construction also when a kept eigenvalue is numerical noise just above eps (the reference's own toy graph has one).
    std::vector<double> Bv((size_t)n * m, 0.0), cv(m, 0.0); // Arm V, V^T bmm
    for (int k = 0; k < m; k++) {
        for (int i = 0; i < n; i++) {
            double s = 0.0;
            for (int j = 0; j < m; j++) s += A[(size_t)(m + i) * N + j] * Vm[(size_t)j * m + k];
            Bv[(size_t)i * m + k] = s;
        }
        double s = 0.0;
        for (int j = 0; j < m; j++) s += Vm[(size_t)j * m + k] * b[j];
        cv[k] = s;
    }
    Ak.assign((size_t)n * n, 0.0);
    bk.assign(n, 0.0);
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) {
            double s = 0.0;
            for (int k = 0; k < m; k++)
                if (wm[k] > eps) s += Bv[(size_t)i * m + k] * (1.0 / wm[k]) * Bv[(size_t)j * m + k];
            Ak[(size_t)i * n + j] = A[(size_t)(m + i) * N + m + j] - s;
        }
        double s = 0.0;
        for (int k = 0; k < m; k++)
            if (wm[k] > eps) s += Bv[(size_t)i * m + k] * (1.0 / wm[k]) * cv[k];
        bk[i] = b[m + i] - s;
    }
    // rank-revealing decomposition of Ak: eigenvalues > eps kept (:318-342)
    std::vector<double> wk, Vk;
    sym_eig(n, Ak.data(), wk, Vk);
    n_full = 0;
    for (int k = 0; k < n; k++) n_full += wk[k] > eps ? 1 : 0;
    U.assign((size_t)n * n_full, 0.0);
    Lambda.assign(n_full, 0.0);
    int q = 0;
    for (int k = 0; k < n; k++) {
        if (!(wk[k] > eps)) continue;
        for (int i = 0; i < n; i++) U[(size_t)i * n_full + q] = Vk[(size_t)i * n + k];
        Lambda[q++] = wk[k];
    }
    // J = Lambda^1/2 U^T, r = -Lambda^-1/2 U^T bk   (:516-530)
    Jm.assign((size_t)n_full * n, 0.0);
    r0.assign(n_full, 0.0);
    for (int k = 0; k < n_full; k++) {
        const double sq = std::sqrt(Lambda[k]), isq = std::sqrt(1.0 / Lambda[k]);
        double s = 0.0;
        for (int i = 0; i < n; i++) {
            Jm[(size_t)k * n + i] = sq * U[(size_t)i * n_full + k];
            s += U[(size_t)i * n_full + k] * bk[i];
        }
        r0[k] = -isq * s;
    }
    return true;
}

} // namespace orc

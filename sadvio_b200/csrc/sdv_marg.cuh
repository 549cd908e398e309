// Marginal-prior construction on the device — replaces, behind sdv_marginalize:
//   Marginalization::computeInformationAndGradient   cpp/src/optimizers/marginalization.cpp:145-211
//   Marginalization::computeSchurComplement          cpp/src/optimizers/marginalization.cpp:213-265
//   Marginalization::rankReveallingDecomposition     cpp/src/optimizers/marginalization.cpp:318-342
//   Marginalization::computeJacobiansAndResiduals    cpp/src/optimizers/marginalization.cpp:516-530
//   Marginalization::sparsifyVIO / sparsifyVO        cpp/src/optimizers/marginalization.cpp:362-514
// driven by AngularAdjustmentCERESAnalytic::marginalize (AngularAdjustmentCERESAnalytic.cpp:488-739) and
// BundleAdjustmentCERESAnalytic::marginalize (BundleAdjustmentCERESAnalytic.cpp:431-660).
//
// The two Eigen::SelfAdjointEigenSolver calls of the reference (m x m and n x n, hundreds of rows) are the dense part.
// Here: BLOCK JACOBI in one-sided storage.  Next to the eigenvector estimate V the kernel keeps G = A V; the block of
// V^T A V that belongs to two column blocks I, J is then V_p^T G_p (p = I u J) and needs no other column.  Columns are
// grouped in blocks of 16; a step pairs the blocks round-robin and ONE CTA owns one pair: it forms that 32 x 32 block,
// diagonalises it with a parallel cyclic Jacobi in shared memory and applies the accumulated rotation to its columns of
// G and V — no CTA ever touches another CTA's columns inside a step, so a step is one launch without any inter-CTA
// synchronisation (a two-sided update A <- J^T A J would have to touch every other CTA's rows).  At the end
// lambda_i = v_i . g_i.  Everything is FP64 on the DFMA pipe (mma.sync.m8n8k4.f64 issues at ~1 / 25 cycles per warp on B200
// and does not scale across warps — measured in round 1, sdv_chol.cuh); the matrices (n <= ~1000: a few MB) stay in L2.
#pragma once
#include "sdv_fused.cuh"
#include "sdv_chol_band.cuh"

namespace sdv {

constexpr int EB = 16;          // columns per block of the block Jacobi
constexpr int EP = 2 * EB;      // panel width of a block pair
constexpr int ET = 512;         // threads per CTA (16 warps: every phase of the Jacobi step is latency-bound, ncu: 20 % issue slots at 8 warps)

// Round-robin pairing ("circle method") of `cnt` players (cnt even): in round `s` (0 .. cnt-2) game k (0 .. cnt/2-1)
SDV_DEV void round_robin(int cnt, int s, int k, int &a, int &b) {
    const int m = cnt - 1;
    if (k == 0) {
        a = m;
        b = s % m;
    } else {
        a = (s + k) % m;
        b = (s + m - k) % m;
    }
}

// device-resident description of one marginalisation (the window itself is the resident DevProblem)
struct MargPlan {
    int N, m, n;
    int col_f0, col_f1;       // first column of frame 0 / frame 1 (-1: no such block)
    int nsel;                 // observations of frame 0 that enter (landmarks kept or marginalised)
    const int *sel_obs;       // [nsel] observation index
    const int *sel_col;       // [nsel] first column of its landmark
    int imu_pair;             // IMU pair (frame 0, frame 1) or -1
    int prior_f0, prior_f1;   // frame index whose PosePriordx enters, or -1
    int nmap;                 // mapped columns of the previous prior (order of DevProblem::mp_src_col)
    const int *mp_col;        // [nmap] their column here, -1 = not a parameter of this marginalisation
};

// ---------------------------------------------------------------------------------------------------------------------
// information matrix A = sum J^T J and gradient b = sum J^T r (marginalization.cpp:145-211) at the current state (dx = 0)
// ---------------------------------------------------------------------------------------------------------------------
// visual factors of frame 0 (…Analytic.cpp:565-629: sigma = 1 / focal; BundleAdjustment…:512-571: sigma = 1)
template <int KIND> __global__ void __launch_bounds__(128) k_marg_visual(const DevProblem *__restrict__ Pg, LinBuf B0, MargPlan M, double *A, double *b) {
    const DevProblem &P = *Pg;
    const int N = M.N, c0 = M.col_f0;
    const int lane = threadIdx.x & 31;
    double H[21], gp[6];
#pragma unroll
    for (int k = 0; k < 21; k++) H[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 6; k++) gp[k] = 0.0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < M.nsel; k += gridDim.x * blockDim.x) {
        const int o = M.sel_obs[k], cl = M.sel_col[k];
        ObsOperands<KIND> q;
        load_obs<KIND>(P, o, q);
        const int c = q.fc % P.C;
        const double w = KIND == 0 ? 0.5 * (P.K[4 * c] + P.K[4 * c + 1]) : 1.0; // 1 / sigma
        const int l = P.obs_lmk[o];
        const double p[3] = {P.lmk_t[3 * (size_t)l], P.lmk_t[3 * (size_t)l + 1], P.lmk_t[3 * (size_t)l + 2]};
        double r[2], Jp[12], Jl[6];
        eval_visual<KIND>(B0.fct + (size_t)q.fc * FCT_ROW, P.K + 4 * c, w, p, q.meas, r, Jp, Jl);
#pragma unroll
        for (int i = 0; i < 6; i++) {
#pragma unroll
            for (int j = 0; j <= i; j++) H[tri_idx(i, j)] += Jp[i] * Jp[j] + Jp[6 + i] * Jp[6 + j];
            gp[i] += Jp[i] * r[0] + Jp[6 + i] * r[1];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const double wv = Jp[i] * Jl[j] + Jp[6 + i] * Jl[3 + j];
                atomicAdd(&A[(size_t)(c0 + i) * N + cl + j], wv);
                atomicAdd(&A[(size_t)(cl + j) * N + c0 + i], wv);
            }
        }
#pragma unroll
        for (int i = 0; i < 3; i++) {
#pragma unroll
            for (int j = 0; j < 3; j++) atomicAdd(&A[(size_t)(cl + i) * N + cl + j], Jl[i] * Jl[j] + Jl[3 + i] * Jl[3 + j]);
            atomicAdd(&b[cl + i], Jl[i] * r[0] + Jl[3 + i] * r[1]);
        }
    }
    // the pose block of frame 0 is shared by every observation: one set of atomics per warp
#pragma unroll
    for (int k = 0; k < 21; k++) H[k] = warp_sum(H[k]);
#pragma unroll
    for (int k = 0; k < 6; k++) gp[k] = warp_sum(gp[k]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 6; i++) {
#pragma unroll
            for (int j = 0; j <= i; j++) {
                atomicAdd(&A[(size_t)(c0 + i) * N + c0 + j], H[tri_idx(i, j)]);
                if (j != i) atomicAdd(&A[(size_t)(c0 + j) * N + c0 + i], H[tri_idx(i, j)]);
            }
            atomicAdd(&b[c0 + i], gp[i]);
        }
    }
}

// IMUFactor + IMUBiasFactor of (frame 0, frame 1) (…Analytic.cpp:510-559), the previous prior as one more block
// (MarginalizationFactor at dx = 0, :631-660), the pose priors (:664-687).  One CTA; the sections touch overlapping
// entries, so they run one after the other.  The factor values come from k_lin_factors at x = 0.
__global__ void __launch_bounds__(256) k_marg_factors(const DevProblem *__restrict__ Pg, LinBuf B0, MargPlan M, double *A, double *b) {
    const DevProblem &P = *Pg;
    const int N = M.N, t = threadIdx.x;
    if (M.imu_pair >= 0) {
        const int p = M.imu_pair, a0 = M.col_f0, a1 = M.col_f1;
        const double *J = B0.imu_J + (size_t)p * 216, *r = B0.imu_r + (size_t)p * 9;
        // parameter blocks pose_i, pose_j, v_i, v_j, ba_i, bg_i (…Analytic.cpp:522-537)
        auto cm = [&](int c) { return c < 6 ? a0 + c : (c < 12 ? a1 + c - 6 : (c < 15 ? a0 + 6 + c - 12 : (c < 18 ? a1 + 6 + c - 15 : a0 + 9 + c - 18))); };
        for (int e = t; e < 24 * 24 + 24; e += blockDim.x) {
            if (e < 576) {
                const int a = e / 24, c = e - a * 24;
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 9; k++) s += J[k * 24 + a] * J[k * 24 + c];
                A[(size_t)cm(a) * N + cm(c)] += s;
            } else {
                const int a = e - 576;
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 9; k++) s += J[k * 24 + a] * r[k];
                b[cm(a)] += s;
            }
        }
        __syncthreads();
        // IMUBiasFactor (residuals.hpp:252-296): r = w (b_j - b_i), blocks ba_i, bg_i, ba_j, bg_j (…Analytic.cpp:549-556)
        if (t < 6) {
            const bool is_ba = t < 3;
            const int k = t % 3;
            const double sig = is_ba ? P.imu_sigma_ba[p] : P.imu_sigma_bg[p];
            const double w = 1.0 / sqrt(P.imu_dt[p] * sig * sig);
            const double rb = B0.bias_r[(size_t)p * 6 + t];
            const int ci = a0 + (is_ba ? 9 : 12) + k, cj = a1 + (is_ba ? 9 : 12) + k;
            A[(size_t)ci * N + ci] += w * w;
            A[(size_t)cj * N + cj] += w * w;
            A[(size_t)ci * N + cj] -= w * w;
            A[(size_t)cj * N + ci] -= w * w;
            b[ci] -= w * rb;
            b[cj] += w * rb;
        }
        __syncthreads();
    }
    if (M.nmap > 0) {
        const int nm = M.nmap;
        for (int e = t; e < nm * nm + nm; e += blockDim.x) {
            if (e < nm * nm) {
                const int a = e / nm, c = e - a * nm;
                if (M.mp_col[a] >= 0 && M.mp_col[c] >= 0) A[(size_t)M.mp_col[a] * N + M.mp_col[c]] += P.mp_H[e];
            } else {
                const int a = e - nm * nm;
                if (M.mp_col[a] >= 0) b[M.mp_col[a]] += P.mp_g0[a];
            }
        }
        __syncthreads();
    }
    for (int q = 0; q < 2; q++) {
        const int f = q == 0 ? M.prior_f0 : M.prior_f1, c0 = q == 0 ? M.col_f0 : M.col_f1;
        if (f < 0) continue;
        const double *J = B0.prior_J + 36 * (size_t)f, *r = B0.prior_r + 6 * (size_t)f;
        if (t < 42) {
            if (t < 36) {
                const int a = t / 6, c = t - a * 6;
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 6; k++) s += J[k * 6 + a] * J[k * 6 + c];
                A[(size_t)(c0 + a) * N + c0 + c] += s;
            } else {
                const int a = t - 36;
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 6; k++) s += J[k * 6 + a] * r[k];
                b[c0 + a] += s;
            }
        }
        __syncthreads();
    }
}

// dst (np x np, zero padded) = 1/2 (src + src^T) of the n x n block of src starting at (off, off)  (marginalization.cpp:228)
__global__ void k_sym_block(const double *src, int lds, int off, int n, double *dst, int np) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)np * np; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / np), j = (int)(e - (size_t)i * np);
        dst[e] = (i < n && j < n) ? 0.5 * (src[(size_t)(off + i) * lds + off + j] + src[(size_t)(off + j) * lds + off + i]) : 0.0;
    }
}
__global__ void k_set_identity(double *V, int np) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)np * np; e += (size_t)gridDim.x * blockDim.x)
        V[e] = (e / np == e % np) ? 1.0 : 0.0;
}

// ---------------------------------------------------------------------------------------------------------------------
// One step of the block Jacobi: CTA `blockIdx.x` owns the block pair round_robin(nbp, step, blockIdx.x).
//   phase 1  H = V_p^T G_p (32 x 32; G = A V, so H is the (I u J) x (I u J) block of V^T A V), 4 x 4 register blocks over
//            32-row tiles staged in shared memory
//   phase 2  parallel cyclic Jacobi on H in shared memory (16 disjoint rotations per step, 31 steps per sweep), rotations
//            accumulated in Q
//   phase 3  panel <- panel Q for G and V: one thread per ROW, the row in registers, Q read as broadcast 16-byte loads
// flags[sweep] collects the largest relative off-diagonal entry |h_pq| / sqrt(|h_pp h_qq|) seen BEFORE rotating; a sweep
// whose maximum stays below `tol` ends the iteration (later launches return at once).  Entries below `floor_rel` x (largest
// column norm of the input, sqrt(*wmax)) are rounding noise of a rank-deficient matrix: left alone, not counted.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int ELD = EP + 2; // row stride of the shared-memory matrices: rows stay 16-byte aligned

__global__ void __launch_bounds__(ET, 1) k_jacobi_pairs(double *G, double *V, int np, int nb, int nbp, int step, unsigned long long *flags, int sweep, double tol,
                                                        int inner_sweeps, const double *wmax, double floor_rel) {
    if (sweep > 0 && __longlong_as_double((long long)flags[sweep - 1]) <= tol) return; // converged in the previous sweep
    int bi, bj;
    round_robin(nbp, step, blockIdx.x, bi, bj);
    if (bi >= nb || bj >= nb) return; // the bye of an odd block count
    if (bi > bj) {
        const int tmp = bi;
        bi = bj;
        bj = tmp;
    }
    __shared__ __align__(16) double W[EP][ELD], Q[EP][ELD], Tv[32][ELD], Tg[32][ELD];
    __shared__ double cs[EB][2];
    __shared__ int pr[EB][2];
    __shared__ double red[ET / 32];
    const int t = threadIdx.x;
    const int ci = EB * bi, cj = EB * bj;
    const double tiny = floor_rel * sqrt(*wmax);
    auto gcol = [&](int c) { return c < EB ? ci + c : cj + c - EB; };
    // ---- phase 1
    {
        constexpr int RG = ET / 64; // row groups: 64 blocks of 4 x 4 entries x RG threads each
        const int rg = t % RG, blk = t / RG, hi = (blk >> 3) * 4, hj = (blk & 7) * 4;
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[a][c] = 0.0;
        for (int r0 = 0; r0 < np; r0 += 32) {
#pragma unroll
            for (int q = 0; q < 1024 / ET; q++) {
                const int e = t + ET * q, rr = e >> 5, cc = e & 31;
                const bool in = r0 + rr < np; // (np is a multiple of 16, the tile has 32 rows)
                const size_t src = (size_t)(r0 + rr) * np + gcol(cc);
                Tv[rr][cc] = in ? V[src] : 0.0;
                Tg[rr][cc] = in ? G[src] : 0.0;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 32 / RG; k++) {
                const int rr = RG * k + rg;
                const double2 v01 = *reinterpret_cast<const double2 *>(&Tv[rr][hi]), v23 = *reinterpret_cast<const double2 *>(&Tv[rr][hi + 2]);
                const double2 g01 = *reinterpret_cast<const double2 *>(&Tg[rr][hj]), g23 = *reinterpret_cast<const double2 *>(&Tg[rr][hj + 2]);
                const double v4[4] = {v01.x, v01.y, v23.x, v23.y}, g4[4] = {g01.x, g01.y, g23.x, g23.y};
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int c = 0; c < 4; c++) acc[a][c] = fma(v4[a], g4[c], acc[a][c]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                double v = acc[a][c];
#pragma unroll
                for (int o = 1; o < RG; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (rg == 0) W[hi + a][hj + c] = v;
            }
    }
    __syncthreads();
    // H is symmetric up to rounding: symmetrise, Q = I, and measure how far from diagonal this block is
    double off = 0.0;
    for (int e = t; e < EP * EP; e += ET) {
        const int i = e >> 5, j = e & 31;
        Q[i][j] = i == j ? 1.0 : 0.0;
        if (i < j) {
            const double h = 0.5 * (W[i][j] + W[j][i]);
            W[i][j] = h;
            W[j][i] = h;
            if (fabs(h) > tiny) {
                const double d = fabs(W[i][i] * W[j][j]);
                off = fmax(off, d > 0.0 ? fabs(h) * rsqrt(d) : 1.0);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) off = fmax(off, __shfl_xor_sync(0xffffffffu, off, o));
    if ((t & 31) == 0) red[t >> 5] = off;
    __syncthreads();
    off = red[0];
#pragma unroll
    for (int q = 1; q < ET / 32; q++) off = fmax(off, red[q]);
    if (t == 0 && off > 0.0) atomicMax(&flags[sweep], (unsigned long long)__double_as_longlong(off));
    if (off <= tol) return; // nothing to rotate (uniform)
    // ---- phase 2: W <- Q^T W Q
    for (int isw = 0; isw < inner_sweeps; isw++) {
        for (int st = 0; st < EP - 1; st++) {
            if (t < EB) {
                int p, q;
                round_robin(EP, st, t, p, q);
                if (p > q) {
                    const int tmp = p;
                    p = q;
                    q = tmp;
                }
                const double app = W[p][p], aqq = W[q][q], apq = W[p][q];
                double c = 1.0, s = 0.0;
                if (fabs(apq) > tiny && apq * apq > 1e-34 * fabs(app * aqq)) {
                    // tan of the Jacobi angle, t = sign(d) 2 apq / (|d| + sqrt(d^2 + 4 apq^2)), from the APPROXIMATE reciprocal /
                    // reciprocal square root units (one MUFU each, ~1e-6 relative): an angle that is slightly off only leaves a
                    // residual the next visit removes, while the rotation itself stays orthogonal to full precision because
                    // c = 1 / sqrt(1 + t^2) is computed exactly for the t that is used.  (FP64 division / sqrt are ~50-instruction
                    // sequences: six of them on the critical path of every one of the 31 steps was most of this phase's time.)
                    const double d = aqq - app, h2 = fma(d, d, 4.0 * apq * apq);
                    double rs, ri;
                    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rs) : "d"(h2));
                    const double den = fabs(d) + h2 * rs;
                    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(ri) : "d"(den));
                    const double tt = (d >= 0.0 ? 2.0 : -2.0) * apq * ri;
                    if (isfinite(tt)) {
                        c = band_rsqrt(fma(tt, tt, 1.0));
                        s = tt * c;
                    }
                }
                pr[t][0] = p;
                pr[t][1] = q;
                cs[t][0] = c;
                cs[t][1] = s;
            }
            __syncthreads();
            for (int e = t; e < 2 * EB * EP; e += ET) { // columns p, q of W and of Q
                const bool isq = e >= EB * EP;
                const int e2 = isq ? e - EB * EP : e, k = e2 >> 5, i = e2 & 31;
                const int p = pr[k][0], q = pr[k][1];
                const double c = cs[k][0], s = cs[k][1];
                double(*Mx)[ELD] = isq ? Q : W;
                const double xp = Mx[i][p], xq = Mx[i][q];
                Mx[i][p] = c * xp - s * xq;
                Mx[i][q] = s * xp + c * xq;
            }
            __syncthreads();
            for (int e = t; e < EB * EP; e += ET) { // rows p, q of W
                const int k = e >> 5, j = e & 31;
                const int p = pr[k][0], q = pr[k][1];
                const double c = cs[k][0], s = cs[k][1];
                const double xp = W[p][j], xq = W[q][j];
                W[p][j] = c * xp - s * xq;
                W[q][j] = s * xp + c * xq;
            }
            __syncthreads();
        }
    }
    // ---- phase 3: panel <- panel Q, for G and for V
    for (int which = 0; which < 2; which++) {
        double *X = which ? V : G;
        for (int r = t; r < np; r += ET) {
            double a[EP];
            double *rowi = X + (size_t)r * np + ci, *rowj = X + (size_t)r * np + cj;
#pragma unroll
            for (int c = 0; c < EB; c += 2) {
                const double2 x = *reinterpret_cast<const double2 *>(rowi + c), y = *reinterpret_cast<const double2 *>(rowj + c);
                a[c] = x.x;
                a[c + 1] = x.y;
                a[EB + c] = y.x;
                a[EB + c + 1] = y.y;
            }
            // the row stays in registers (static indices); the output columns are walked two at a time in a ROLLED loop so that
            // ptxas cannot hoist the 512 broadcast loads of Q in front of the arithmetic (it spilled 7 KB per thread when it did)
#pragma unroll
            for (int half = 0; half < 2; half++) {
                double *dst = half ? rowj : rowi;
#pragma unroll 1
                for (int j = 0; j < EB; j += 2) {
                    double o0a = 0.0, o0b = 0.0, o1a = 0.0, o1b = 0.0;
#pragma unroll
                    for (int i = 0; i < EP; i += 2) {
                        const double2 qa = *reinterpret_cast<const double2 *>(&Q[i][EB * half + j]); // same address in every lane: broadcast
                        const double2 qb = *reinterpret_cast<const double2 *>(&Q[i + 1][EB * half + j]);
                        o0a = fma(a[i], qa.x, o0a);
                        o1a = fma(a[i], qa.y, o1a);
                        o0b = fma(a[i + 1], qb.x, o0b);
                        o1b = fma(a[i + 1], qb.y, o1b);
                    }
                    *reinterpret_cast<double2 *>(dst + j) = make_double2(o0a + o0b, o1a + o1b);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same Jacobi step with a THREAD-BLOCK CLUSTER per block pair (default; k_jacobi_pairs above remains the one-CTA fallback).
// With ~20 pairs per step the one-CTA kernel keeps 20 of 148 SMs busy and walks all np rows twice per step (phase 1 and
// phase 3: ~2/3 of its time).  Here CS CTAs share a pair: CTA r owns the rows [r rows_per_cta, (r + 1) rows_per_cta) of G and V,
//   phase 1  forms its partial of the 32 x 32 block over its rows, the partials are summed through distributed shared memory
//            IN RANK ORDER by every CTA, so all of them hold the bit-identical block and take identical decisions;
//   phase 2  every CTA runs the same small Jacobi redundantly (nothing to exchange).  W <- J^T W J is ONE in-place pass per step (one
//            thread per 2 x 2 group of a row pair and a column pair), Q <- Q J beside it: two CTA barriers per step instead of three;
//   phase 3  panel <- panel Q on its own rows only, one thread per (matrix, half panel, row): 4 x 128 work items per pass.
// No CTA touches another CTA's rows or another pair's columns: still one launch per step, no global synchronisation.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ET, 1) k_jacobi_pairs_cl(double *G, double *V, int np, int nb, int nbp, int step, unsigned long long *flags, int sweep, double tol,
                                                           int inner_sweeps, const double *wmax, double floor_rel, int rows_per_cta) {
    if (sweep > 0 && __longlong_as_double((long long)flags[sweep - 1]) <= tol) return; // converged in the previous sweep (uniform over the grid)
    const int CS = (int)cluster_size(), rank = (int)cluster_rank();
    int bi, bj;
    round_robin(nbp, step, (int)blockIdx.x / CS, bi, bj);
    if (bi >= nb || bj >= nb) return; // the bye of an odd block count (uniform over the cluster)
    if (bi > bj) {
        const int tmp = bi;
        bi = bj;
        bj = tmp;
    }
    __shared__ __align__(16) double W[EP][ELD], Q[EP][ELD], Tv[32][ELD], Tg[32][ELD];
    __shared__ double ca[EB], cb[EB]; // cos / sin of the 16 rotations of a step
    __shared__ int pr[EB][2];
    __shared__ double red[ET / 32];
    const int t = threadIdx.x;
    const int ci = EB * bi, cj = EB * bj;
    const double tiny = floor_rel * sqrt(*wmax);
    auto gcol = [&](int c) { return c < EB ? ci + c : cj + c - EB; };
    const int r_lo = rank * rows_per_cta, r_hi = min(np, r_lo + rows_per_cta);
    // ---- phase 1: partial block over this CTA's rows -> Tg
    {
        constexpr int RG = ET / 64;
        const int rg = t % RG, blk = t / RG, hi = (blk >> 3) * 4, hj = (blk & 7) * 4;
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[a][c] = 0.0;
        for (int r0 = r_lo; r0 < r_hi; r0 += 32) {
#pragma unroll
            for (int q = 0; q < 1024 / ET; q++) {
                const int e = t + ET * q, rr = e >> 5, cc = e & 31;
                const bool in = r0 + rr < r_hi;
                const size_t src = (size_t)(r0 + rr) * np + gcol(cc);
                Tv[rr][cc] = in ? V[src] : 0.0;
                Tg[rr][cc] = in ? G[src] : 0.0;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 32 / RG; k++) {
                const int rr = RG * k + rg;
                const double2 v01 = *reinterpret_cast<const double2 *>(&Tv[rr][hi]), v23 = *reinterpret_cast<const double2 *>(&Tv[rr][hi + 2]);
                const double2 g01 = *reinterpret_cast<const double2 *>(&Tg[rr][hj]), g23 = *reinterpret_cast<const double2 *>(&Tg[rr][hj + 2]);
                const double v4[4] = {v01.x, v01.y, v23.x, v23.y}, g4[4] = {g01.x, g01.y, g23.x, g23.y};
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int c = 0; c < 4; c++) acc[a][c] = fma(v4[a], g4[c], acc[a][c]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                double v = acc[a][c];
#pragma unroll
                for (int o = 1; o < RG; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (rg == 0) Tg[hi + a][hj + c] = v;
            }
    }
    cluster_sync_all(); // every partial is in its CTA's Tg (barrier.cluster: release / acquire, all threads of all CTAs)
    for (int e = t; e < EP * EP; e += ET) {
        const int i = e >> 5, j = e & 31;
        double sum = 0.0;
        for (int r = 0; r < CS; r++) sum += dsmem_load(&Tg[i][j], (unsigned)r); // rank order: the same sum in every CTA
        W[i][j] = sum;
    }
    cluster_sync_all(); // nobody reads this CTA's Tg any more (it may exit or reuse it)
    // H is symmetric up to rounding: symmetrise, Q = I, and measure how far from diagonal this block is
    double off = 0.0;
    for (int e = t; e < EP * EP; e += ET) {
        const int i = e >> 5, j = e & 31;
        Q[i][j] = i == j ? 1.0 : 0.0;
        if (i < j) {
            const double h = 0.5 * (W[i][j] + W[j][i]);
            W[i][j] = h;
            W[j][i] = h;
            if (fabs(h) > tiny) {
                const double d = fabs(W[i][i] * W[j][j]);
                off = fmax(off, d > 0.0 ? fabs(h) * rsqrt(d) : 1.0);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) off = fmax(off, __shfl_xor_sync(0xffffffffu, off, o));
    if ((t & 31) == 0) red[t >> 5] = off;
    __syncthreads();
    off = red[0];
#pragma unroll
    for (int q = 1; q < ET / 32; q++) off = fmax(off, red[q]);
    if (t == 0 && rank == 0 && off > 0.0) atomicMax(&flags[sweep], (unsigned long long)__double_as_longlong(off));
    if (off <= tol) return; // nothing to rotate (uniform over the cluster: identical W)
    // ---- phase 2: W <- Q^T W Q, redundantly in every CTA of the cluster.  The phase is bound by shared-memory traffic, so a rotation
    // step touches every entry once: thread (k, m) < 256 owns the 2 x 2 group {p_k, q_k} x {p_m, q_m} of W (both rotations applied to
    // it in registers, in place: 4 loads + 4 stores for 4 entries), threads >= 256 rotate the columns of Q.
    for (int isw = 0; isw < inner_sweeps; isw++) {
        for (int st = 0; st < EP - 1; st++) {
            if (t < EB) {
                int p, q;
                round_robin(EP, st, t, p, q);
                if (p > q) {
                    const int tmp = p;
                    p = q;
                    q = tmp;
                }
                const double app = W[p][p], aqq = W[q][q], apq = W[p][q];
                double c = 1.0, s = 0.0;
                if (fabs(apq) > tiny && apq * apq > 1e-34 * fabs(app * aqq)) { // (angle from the approximate units, exact c for the t used: see k_jacobi_pairs)
                    const double d = aqq - app, h2 = fma(d, d, 4.0 * apq * apq);
                    double rs, ri;
                    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rs) : "d"(h2));
                    const double den = fabs(d) + h2 * rs;
                    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(ri) : "d"(den));
                    const double tt = (d >= 0.0 ? 2.0 : -2.0) * apq * ri;
                    if (isfinite(tt)) {
                        c = band_rsqrt(fma(tt, tt, 1.0));
                        s = tt * c;
                    }
                }
                pr[t][0] = p;
                pr[t][1] = q;
                ca[t] = c; // x_p' = c x_p - s x_q,  x_q' = s x_p + c x_q
                cb[t] = s;
            }
            __syncthreads();
            if (t < EB * EB) {
                const int kk = t >> 4, mm = t & 15;
                const int p = pr[kk][0], q = pr[kk][1], r = pr[mm][0], u = pr[mm][1];
                const double ck = ca[kk], sk = cb[kk], cm = ca[mm], sm = cb[mm];
                const double w00 = W[p][r], w01 = W[p][u], w10 = W[q][r], w11 = W[q][u];
                const double a1 = cm * w00 - sm * w01, b1 = sm * w00 + cm * w01, c1 = cm * w10 - sm * w11, d1 = sm * w10 + cm * w11; // columns r, u
                W[p][r] = ck * a1 - sk * c1; // rows p, q
                W[q][r] = sk * a1 + ck * c1;
                W[p][u] = ck * b1 - sk * d1;
                W[q][u] = sk * b1 + ck * d1;
            } else {
                for (int e = t - EB * EB; e < EB * EP; e += ET - EB * EB) { // columns p, q of Q (one thread owns both entries of a row)
                    const int kk = e >> 5, i = e & 31;
                    const int p = pr[kk][0], q = pr[kk][1];
                    const double c = ca[kk], s = cb[kk];
                    const double xp = Q[i][p], xq = Q[i][q];
                    Q[i][p] = c * xp - s * xq;
                    Q[i][q] = s * xp + c * xq;
                }
            }
            __syncthreads();
        }
    }
    // ---- phase 3: panel <- panel Q for G and V on this CTA's rows; thread = (matrix, half panel, row), 128 rows per pass
    {
        const int wh = t >> 7, which = wh >> 1, half = wh & 1;
        double *X = which ? V : G;
        for (int cbase = r_lo; cbase < r_hi; cbase += 128) {
            const int r = cbase + (t & 127);
            const bool valid = r < r_hi;
            double a[EP];
            double *rowi = X + (size_t)r * np + ci, *rowj = X + (size_t)r * np + cj;
            if (valid) {
#pragma unroll
                for (int c = 0; c < EB; c += 2) {
                    const double2 x = *reinterpret_cast<const double2 *>(rowi + c), y = *reinterpret_cast<const double2 *>(rowj + c);
                    a[c] = x.x;
                    a[c + 1] = x.y;
                    a[EB + c] = y.x;
                    a[EB + c + 1] = y.y;
                }
            }
            __syncthreads(); // both halves of a row have been read before either is overwritten
            if (valid) {
                double *dst = half ? rowj : rowi;
#pragma unroll 1
                for (int j = 0; j < EB; j += 2) { // (rolled: see k_jacobi_pairs)
                    double o0a = 0.0, o0b = 0.0, o1a = 0.0, o1b = 0.0;
#pragma unroll
                    for (int i = 0; i < EP; i += 2) {
                        const double2 qa = *reinterpret_cast<const double2 *>(&Q[i][EB * half + j]); // same address in every lane of the warp: broadcast
                        const double2 qb = *reinterpret_cast<const double2 *>(&Q[i + 1][EB * half + j]);
                        o0a = fma(a[i], qa.x, o0a);
                        o1a = fma(a[i], qa.y, o1a);
                        o0b = fma(a[i + 1], qb.x, o0b);
                        o1b = fma(a[i + 1], qb.y, o1b);
                    }
                    *reinterpret_cast<double2 *>(dst + j) = make_double2(o0a + o0b, o1a + o1b);
                }
            }
        }
    }
}

// largest squared column norm of the (symmetric) input matrix
__global__ void k_max_colnorm2(const double *G, int np, double *out) {
    __shared__ double red[8];
    double m = 0.0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < np; r++) s = fma(G[(size_t)r * np + i], G[(size_t)r * np + i], s);
        m = fmax(m, s);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < (int)(blockDim.x >> 5); q++) m = fmax(m, red[q]);
        *out = m;
    }
}

// lambda_i = v_i . g_i  (G = A V with orthogonal columns: g_i = lambda_i v_i)
__global__ void k_eig_values(const double *G, const double *V, int np, double *w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    double s = 0.0;
    for (int r = 0; r < np; r++) s = fma(V[(size_t)r * np + i], G[(size_t)r * np + i], s);
    w[i] = s;
}

// winv[k] = 1 / w[k] for w[k] > eps, else 0   (the pseudo-inverse of marginalization.cpp:234-240)
__global__ void k_pinv_diag(const double *w, int n, double eps, double *winv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) winv[i] = w[i] > eps ? 1.0 / w[i] : 0.0;
}

// C[i][j] = alpha * sum_k A[i][k] d[k] B(k, j) + beta * Cin[i][j];  B(k, j) = transB ? B[j][k] : B[k][j];  d may be null.
// 32 x 32 tile per CTA of 16 x 16 threads, 2 x 2 outputs each.  The products of a marginalisation are a few hundred rows.
__global__ void __launch_bounds__(256) k_mm(double *C, int ldc, const double *A, int lda, const double *B, int ldb, int transB, const double *d, int Mr, int Nc, int K,
                                            double alpha, const double *Cin, int ldcin, double beta) {
    __shared__ double As[32][33], Bs[32][33];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    double acc[2][2] = {{0, 0}, {0, 0}};
    for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int e = threadIdx.x + 256 * q, rr = e >> 5, cc = e & 31;
            const int i = i0 + rr, k = k0 + cc;
            As[rr][cc] = (i < Mr && k < K) ? A[(size_t)i * lda + k] * (d ? d[k] : 1.0) : 0.0;
            // Bs[kk][jj]
            if (transB) {
                const int j = j0 + rr, kb = k0 + cc;
                Bs[cc][rr] = (j < Nc && kb < K) ? B[(size_t)j * ldb + kb] : 0.0;
            } else {
                const int kb = k0 + rr, j = j0 + cc;
                Bs[rr][cc] = (kb < K && j < Nc) ? B[(size_t)kb * ldb + j] : 0.0;
            }
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < 32; kk++) {
            const double a0 = As[ty][kk], a1 = As[ty + 16][kk], b0 = Bs[kk][tx], b1 = Bs[kk][tx + 16];
            acc[0][0] = fma(a0, b0, acc[0][0]);
            acc[0][1] = fma(a0, b1, acc[0][1]);
            acc[1][0] = fma(a1, b0, acc[1][0]);
            acc[1][1] = fma(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int i = i0 + ty + 16 * a, j = j0 + tx + 16 * c;
            if (i < Mr && j < Nc) C[(size_t)i * ldc + j] = alpha * acc[a][c] + (Cin ? beta * Cin[(size_t)i * ldcin + j] : 0.0);
        }
}

// J = Lambda^1/2 U^T, r0 = -Lambda^-1/2 U^T bk (marginalization.cpp:516-530); U = the eigenvectors `order` selects (> eps, ascending)
__global__ void k_marg_build(const double *V, int np, const double *w, const int *order, int n, int n_full, const double *bk, double *J, double *r0, double *U,
                             double *Lambda) {
    const int k = blockIdx.x; // one CTA per kept eigenpair
    if (k >= n_full) return;
    const int c = order[k];
    const double lam = w[c], sq = sqrt(lam), isq = sqrt(1.0 / lam);
    __shared__ double red[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double u = V[(size_t)i * np + c];
        J[(size_t)k * n + i] = sq * u;
        U[(size_t)i * n_full + k] = u;
        s = fma(u, bk[i], s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int q = 0; q < (int)(blockDim.x >> 5); q++) tot += red[q];
        r0[k] = -isq * tot;
        Lambda[k] = lam;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// small dense helpers for the sparsification (3 x 3 and 15 x 15 symmetric problems, one warp each)
// ---------------------------------------------------------------------------------------------------------------------
// cyclic Jacobi on the n x n symmetric matrix A (shared memory, leading dimension ld), eigenvectors in Vm; lanes update
// the rows / columns of a rotation in parallel.  A is overwritten by diag(w).
SDV_DEV void warp_sym_eig(double *A, double *Vm, int n, int ld) {
    const int lane = threadIdx.x & 31;
    for (int e = lane; e < n * n; e += 32) Vm[(e / n) * ld + e % n] = (e / n == e % n) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 30; sweep++) {
        double offn = 0.0, diag = 0.0;
        for (int e = lane; e < n * n; e += 32) {
            const double v = A[(e / n) * ld + e % n];
            if (e / n == e % n) diag += v * v;
            else offn += v * v;
        }
        offn = warp_sum(offn);
        diag = warp_sum(diag);
        offn = __shfl_sync(0xffffffffu, offn, 0);
        diag = __shfl_sync(0xffffffffu, diag, 0);
        if (offn <= 1e-30 * (diag + 1e-300)) break;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = A[p * ld + q];
                if (apq == 0.0) continue; // uniform: every lane reads the same value
                const double app = A[p * ld + p], aqq = A[q * ld + q];
                const double theta = (aqq - app) / (2.0 * apq);
                const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
                __syncwarp();
                if (lane < n) {
                    const double akp = A[lane * ld + p], akq = A[lane * ld + q];
                    A[lane * ld + p] = c * akp - s * akq;
                    A[lane * ld + q] = s * akp + c * akq;
                    const double vkp = Vm[lane * ld + p], vkq = Vm[lane * ld + q];
                    Vm[lane * ld + p] = c * vkp - s * vkq;
                    Vm[lane * ld + q] = s * vkp + c * vkq;
                }
                __syncwarp();
                if (lane < n) {
                    const double apk = A[p * ld + lane], aqk = A[q * ld + lane];
                    A[p * ld + lane] = c * apk - s * aqk;
                    A[q * ld + lane] = s * apk + c * aqk;
                }
                __syncwarp();
            }
    }
    __syncwarp();
}

// out (n x n, row-major, leading dimension n) = Vm diag(f(w)) Vm^T with f chosen by `mode`:
//   0: inf = M^-1, then the SelfAdjointEigenSolver square root of inf with its eigenvalues <= eps dropped (sparsifyVIO,
//      marginalization.cpp:377-385, :398-406):  f = sqrt(1 / w) if 1 / w > eps else 0
//   1: pseudo-inverse square root of M itself (sparsifyVO, :481-487, :501-507):  f = sqrt(1 / w) if w > eps else 0
SDV_DEV void warp_sqrt_info(const double *A, const double *Vm, int n, int ld, double eps, int mode, double *out) {
    const int lane = threadIdx.x & 31;
    for (int e = lane; e < n * n; e += 32) {
        const int i = e / n, j = e % n;
        double s = 0.0;
        for (int k = 0; k < n; k++) {
            const double w = A[k * ld + k];
            double f;
            if (mode == 0) {
                const double iw = 1.0 / w;
                f = iw > eps ? sqrt(iw) : 0.0;
            } else {
                f = w > eps ? sqrt(1.0 / w) : 0.0;
            }
            s += Vm[i * ld + k] * f * Vm[j * ld + k];
        }
        out[e] = s;
    }
}

// rows of Jt = J U for the measurement functions of the sparsified factors, contracted with Sigma = 1 / Lambda:
// M[a][c] = sum_q Jt[a][q] Jt[c][q] / Lambda[q]; rowfn(a, q) returns Jt[a][q]
template <class RowFn> SDV_DEV void warp_gram_sigma(int rows, int n_full, const double *Lambda, RowFn rowfn, double *Mx, int ld) {
    const int lane = threadIdx.x & 31;
    for (int e = 0; e < rows * rows; e++) {
        const int a = e / rows, c = e % rows;
        if (c < a) continue;
        double s = 0.0;
        for (int q = lane; q < n_full; q += 32) s += rowfn(a, q) * rowfn(c, q) / Lambda[q];
        s = warp_sum(s);
        s = __shfl_sync(0xffffffffu, s, 0);
        if (lane == 0) {
            Mx[a * ld + c] = s;
            Mx[c * ld + a] = s;
        }
    }
    __syncwarp();
}

// sparsifyVIO (marginalization.cpp:362-411): warp 0 of CTA 0 = the absolute factor on the kept frame (IMUPriordx), then one
// warp per kept landmark = its PoseToLandmarkFactor.  U is n x n_full; the kept frame sits at columns 0..14, landmark k at 15 + 3k.
__global__ void __launch_bounds__(32) k_sparsify_vio(const DevProblem *__restrict__ Pg, int f1, const int *keep_lmk, int n_keep, const double *U, const double *Lambda, int n_full,
                                                     double eps, double *imu_sqrt_inf, double *p2l_delta, double *p2l_sqrt_inf) {
    const DevProblem &P = *Pg;
    __shared__ double Mx[15 * 16], Vm[15 * 16];
    const int lane = threadIdx.x;
    const double *T = P.T_f_w + 12 * (size_t)f1;
    const double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]}, tt[3] = {T[3], T[7], T[11]};
    auto u = [&](int i, int q) { return U[(size_t)i * n_full + q]; };
    if (blockIdx.x == 0) {
        // J (15 x n): identity on the frame's columns with J[0:3,0:3] = R, J[0:3,3:6] = R, J[3:6,3:6] = R (:393-396, as written)
        auto row = [&](int a, int q) {
            if (a < 3) return R[a * 3] * (u(0, q) + u(3, q)) + R[a * 3 + 1] * (u(1, q) + u(4, q)) + R[a * 3 + 2] * (u(2, q) + u(5, q));
            if (a < 6) return R[(a - 3) * 3] * u(3, q) + R[(a - 3) * 3 + 1] * u(4, q) + R[(a - 3) * 3 + 2] * u(5, q);
            return u(a, q);
        };
        warp_gram_sigma(15, n_full, Lambda, row, Mx, 16);
        warp_sym_eig(Mx, Vm, 15, 16);
        warp_sqrt_info(Mx, Vm, 15, 16, eps, 0, imu_sqrt_inf);
        return;
    }
    const int k = blockIdx.x - 1;
    if (k >= n_keep) return;
    const int l = keep_lmk[k], c = 15 + 3 * k;
    // J (3 x n): R at the landmark's columns, -R [t]x at 0:3, R at 3:6 (:372-374)
    const double S[9] = {0, -tt[2], tt[1], tt[2], 0, -tt[0], -tt[1], tt[0], 0};
    double RS[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) RS[i * 3 + j] = -(R[i * 3] * S[j] + R[i * 3 + 1] * S[3 + j] + R[i * 3 + 2] * S[6 + j]);
    auto row = [&](int a, int q) {
        double s = 0.0;
        for (int j = 0; j < 3; j++) s += R[a * 3 + j] * (u(c + j, q) + u(3 + j, q)) + RS[a * 3 + j] * u(j, q);
        return s;
    };
    warp_gram_sigma(3, n_full, Lambda, row, Mx, 16);
    warp_sym_eig(Mx, Vm, 3, 16);
    warp_sqrt_info(Mx, Vm, 3, 16, eps, 0, p2l_sqrt_inf + 9 * (size_t)k);
    if (lane < 3) // t_f_lmk (:387)
        p2l_delta[3 * (size_t)k + lane] = R[lane * 3] * P.lmk_t[3 * (size_t)l] + R[lane * 3 + 1] * P.lmk_t[3 * (size_t)l + 1] + R[lane * 3 + 2] * P.lmk_t[3 * (size_t)l + 2] + tt[lane];
}

// sparsifyVO, first half (marginalization.cpp:420-431, :267-274): coupling |trace(Ak_kl)| of every pair of kept landmarks and the
// entropy of every kept landmark under the marginal covariance U Sigma U^T.  One warp per kept landmark.
__global__ void __launch_bounds__(32) k_vo_coupling(const double *Ak, int n, int first, int K, const double *U, const double *Lambda, int n_full, double *mi, double *ent) {
    const int k = blockIdx.x, lane = threadIdx.x;
    if (k >= K) return;
    const int ck = first + 3 * k;
    for (int l = lane; l < K; l += 32) {
        const int cl = first + 3 * l;
        mi[(size_t)k * K + l] = k == l ? 0.0 : fabs(Ak[(size_t)ck * n + cl] + Ak[(size_t)(ck + 1) * n + cl + 1] + Ak[(size_t)(ck + 2) * n + cl + 2]);
    }
    double S[6] = {0, 0, 0, 0, 0, 0};
    for (int q = lane; q < n_full; q += 32) {
        const double u0 = U[(size_t)ck * n_full + q], u1 = U[(size_t)(ck + 1) * n_full + q], u2 = U[(size_t)(ck + 2) * n_full + q], sg = 1.0 / Lambda[q];
        S[0] += u0 * u0 * sg; S[1] += u0 * u1 * sg; S[2] += u0 * u2 * sg; S[3] += u1 * u1 * sg; S[4] += u1 * u2 * sg; S[5] += u2 * u2 * sg;
    }
    for (int i = 0; i < 6; i++) S[i] = warp_sum(S[i]);
    if (lane == 0) {
        const double det = S[0] * (S[3] * S[5] - S[4] * S[4]) - S[1] * (S[1] * S[5] - S[4] * S[2]) + S[2] * (S[1] * S[4] - S[3] * S[2]);
        // std::pow(2 pi e, size / 2) with the INTEGER division the reference writes (3 / 2 = 1), marginalization.cpp:273
        ent[k] = log(2.0 * 3.14159265358979323846 * 2.71828182845904523536 * det);
    }
}

// sparsifyVO, second half (:474-511): the unary Landmark3DPrior of `with_prior` (CTA 0) and one LandmarkToLandmarkFactor per
// consecutive pair of the chain.  chain_col[k] = first column of the k-th landmark of the chain.
__global__ void __launch_bounds__(32) k_sparsify_vo(const DevProblem *__restrict__ Pg, int col_prior, const int *chain_col, const int *chain_lmk, int n_chain, const double *U,
                                                    const double *Lambda, int n_full, double eps, double *lmk_sqrt_inf, double *l2l_delta, double *l2l_sqrt_inf) {
    const DevProblem &P = *Pg;
    __shared__ double Mx[3 * 4], Vm[3 * 4];
    const int lane = threadIdx.x;
    auto u = [&](int i, int q) { return U[(size_t)i * n_full + q]; };
    if (blockIdx.x == 0) {
        auto row = [&](int a, int q) { return u(col_prior + a, q); };
        warp_gram_sigma(3, n_full, Lambda, row, Mx, 4);
        warp_sym_eig(Mx, Vm, 3, 4);
        warp_sqrt_info(Mx, Vm, 3, 4, eps, 1, lmk_sqrt_inf);
        return;
    }
    const int k = blockIdx.x - 1;
    if (k + 1 >= n_chain) return;
    const int ca = chain_col[k], cb = chain_col[k + 1];
    auto row = [&](int a, int q) { return u(ca + a, q) - u(cb + a, q); };
    warp_gram_sigma(3, n_full, Lambda, row, Mx, 4);
    warp_sym_eig(Mx, Vm, 3, 4);
    warp_sqrt_info(Mx, Vm, 3, 4, eps, 1, l2l_sqrt_inf + 9 * (size_t)k);
    if (lane < 3) l2l_delta[3 * (size_t)k + lane] = P.lmk_t[3 * (size_t)chain_lmk[k] + lane] - P.lmk_t[3 * (size_t)chain_lmk[k + 1] + lane];
}

} // namespace sdv

// AOptimizer::VIInit on the device (the fourth "other" solve): visual-inertial initialisation of the
// gravity direction, the keyframe velocities and — optionally — the metric scale of an up-to-scale visual trajectory
// (reference cpp/src/optimizers/AOptimizer.cpp:448-581, functor IMUFactorInit cpp/include/isaeslam/optimizers/residuals.hpp:302-410).
//
// The problem is tiny and dense-ish (one 3-vector per keyframe, chained by the IMU pairs, bordered by the 2-dof gravity
// alignment and the log-scale: n = 3F + 3), and it runs once per session.  The whole Ceres LM loop therefore lives in ONE
// kernel on ONE CTA: no launch, graph or host round trip per iteration.  Per iteration: one thread per IMU pair evaluates the
// whitened 9x9 Jacobian block and residual (the functor's rows 0-2 — the rotation residual — do not depend on any free
// parameter, dba / dbg being constant at zero, AOptimizer.cpp:472-477, but they count in the cost), the pairs are summed into
// the normal equations in a fixed order (deterministic), a right-looking Cholesky in shared memory solves the damped system,
// the candidate is evaluated residual-only, and thread 0 runs the same accept / reject / tolerance logic as the window solve
// (ctrl_step, sdv_kernels.cuh).  Jacobi scaling and the LM diagonal follow DESIGN.md section 3 (unscaled variables).
//
// Quirk kept on purpose: the functor's derivative with respect to the log-scale omits the factor exp(lambda)
// (residuals.hpp:399-400); Ceres uses the Jacobian it is given, so does this kernel.
#pragma once
#include "sdv_kernels.cuh"

namespace sdv {

struct VIInitArgs {
    int F, P, n, optim_scale, max_iter, a_in_smem;
    const double *T_f_w, *v;                   // [F][12], [F][3]
    const int *imu_i, *imu_j;                  // [P]
    const double *imu_dt, *imu_dR, *imu_dv, *imu_dp; // [P], [P][9], [P][3], [P][3]
    const double *inf_sqrt;                    // [P][81] upper-triangular U, U^T U = cov^-1 (k_imu_inf_sqrt)
    double *Jw, *rw;                           // [P][81], [P][9]: whitened Jacobian block (9 x [w0 w1 | vi | vj | lambda]) and residual
    double *A;                                 // [n][n] normal equations / factor when they do not fit in shared memory
    double *vecs;                              // [8][n]: x0, x1, g, hdiag, scale, damp, delta, y
    double *out;                               // [3F + 3]: dv, r_wi, lambda
    LMState *st;
    Accum *acc;
    SolverOpts opt;
};

constexpr int VI_THREADS = 256;

// One IMUFactorInit::Evaluate at the parameter vector x (layout: v blocks 3f.., r_wi at 3F, lambda at 3F+2).  Returns the squared
// norm of the whitened residual; Jw (may be null) receives the whitened 9x9 block, columns [w0 w1 | dv_i | dv_j | lambda].
SDV_DEV double viinit_eval(const VIInitArgs &a, int p, const double *x, double *rw_out, double *Jw_out) {
    const int i = a.imu_i[p], j = a.imu_j[p], F = a.F;
    const double dt = a.imu_dt[p];
    const double g[3] = {0.0, 0.0, -9.81}; // data/sensors/IMU.h:8
    double Ri[9], ti[3], Rj[9], tj[3];
    load_RT(a.T_f_w + 12 * i, Ri, ti);
    load_RT(a.T_f_w + 12 * j, Rj, tj);
    const double w[3] = {x[3 * F], x[3 * F + 1], 0.0};                        // residuals.hpp:309
    double Rwi[9], RiRw[9];
    exp_so3(w, Rwi);
    mat3_mul(Ri, Rwi, RiRw);
    const double lam = a.optim_scale ? x[3 * F + 2] : 0.0, el = exp(lam);
    double vi[3], vj[3], pi[3], pj[3], av[3], ap[3], dpos[3];
    matT3_vec(Ri, ti, pi);                                                    // T.inverse().translation() = -R^T t
    matT3_vec(Rj, tj, pj);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        vi[k] = a.v[3 * i + k] + x[3 * i + k];                                // :311-312
        vj[k] = a.v[3 * j + k] + x[3 * j + k];
        dpos[k] = -pj[k] + pi[k];
        av[k] = (vj[k] - vi[k]) - g[k] * dt;
        ap[k] = el * dpos[k] - vi[k] * dt - 0.5 * g[k] * dt * dt;
    }
    double e[9];
    {   // r_dr = Log(delta_R^T R_i R_j^T)  (d_bg = 0: Exp(J_dR_bg d_bg) = I), :326-328
        double RiRjT[9], dR[9];
        mat3_mulT(Ri, Rj, RiRjT);
        matT3_mul(a.imu_dR + 9 * p, RiRjT, dR);
        log_so3(dR, e);
    }
    double t3[3];
    mat3_vec(RiRw, av, t3);
#pragma unroll
    for (int k = 0; k < 3; k++) e[3 + k] = t3[k] - a.imu_dv[3 * p + k];      // :329-330
    mat3_vec(RiRw, ap, t3);
#pragma unroll
    for (int k = 0; k < 3; k++) e[6 + k] = t3[k] - a.imu_dp[3 * p + k];      // :331-335
    const double *U = a.inf_sqrt + 81 * (size_t)p;
    double c = 0.0;
#pragma unroll
    for (int r = 0; r < 9; r++) {
        double s = 0.0;
        for (int q = r; q < 9; q++) s += U[r * 9 + q] * e[q];                 // :340 (U is upper triangular)
        rw_out[r] = s;
        c += s * s;
    }
    if (Jw_out) {
        // unwhitened rows 3..8 of the 9x9 block; rows 0..2 are zero
        double J[6][9];
        double Jr[9], S[9], T1[9], Bv[9], Bp[9];
        right_jacobian(w, Jr);
        skew3(av, S);
        mat3_mul(RiRw, S, T1);
        mat3_mul(T1, Jr, Bv);                                                 // :348-349 (sign below)
        skew3(ap, S);
        mat3_mul(RiRw, S, T1);
        mat3_mul(T1, Jr, Bp);                                                 // :350-354
        double sc[3];
        mat3_vec(RiRw, dpos, sc);                                             // :399-400 (no exp(lambda) factor)
#pragma unroll
        for (int r = 0; r < 3; r++) {
            J[r][0] = -Bv[r * 3 + 0];
            J[r][1] = -Bv[r * 3 + 1];
            J[3 + r][0] = -Bp[r * 3 + 0];
            J[3 + r][1] = -Bp[r * 3 + 1];
#pragma unroll
            for (int cc = 0; cc < 3; cc++) {
                J[r][2 + cc] = -RiRw[r * 3 + cc];                             // :362
                J[3 + r][2 + cc] = -RiRw[r * 3 + cc] * dt;                    // :363
                J[r][5 + cc] = RiRw[r * 3 + cc];                              // :371
                J[3 + r][5 + cc] = 0.0;
            }
            J[r][8] = 0.0;
            J[3 + r][8] = sc[r];
        }
        for (int r = 0; r < 9; r++)
#pragma unroll
            for (int cc = 0; cc < 9; cc++) {
                double s = 0.0;
                for (int q = (r > 3 ? r : 3); q < 9; q++) s += U[r * 9 + q] * J[q - 3][cc];
                Jw_out[r * 9 + cc] = s;
            }
    }
    return c;
}

SDV_DEV int viinit_col(const VIInitArgs &a, int p, int c) { // column of entry c of pair p's block in the normal equations, -1 = constant
    if (c < 2) return 3 * a.F + c;
    if (c < 5) return 3 * a.imu_i[p] + (c - 2);
    if (c < 8) return 3 * a.imu_j[p] + (c - 5);
    return a.optim_scale ? 3 * a.F + 2 : -1;
}

// block-wide sum of one double per thread (all threads get the result)
SDV_DEV double viinit_block_sum(double v, double *red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int k = 0; k < VI_THREADS / 32; k++) s += red[k];
    return s;
}

__global__ void __launch_bounds__(VI_THREADS) k_viinit(VIInitArgs a) {
    extern __shared__ double vi_smem[];
    __shared__ LMState st;
    __shared__ Accum acc;
    __shared__ double red[VI_THREADS / 32];
    __shared__ int fail_s;
    const int tid = threadIdx.x, n = a.n, P = a.P;
    double *A = a.a_in_smem ? vi_smem : a.A;
    double *xb[2] = {a.vecs, a.vecs + n};
    double *g = a.vecs + 2 * n, *hd = a.vecs + 3 * n, *scale = a.vecs + 4 * n, *damp = a.vecs + 5 * n, *delta = a.vecs + 6 * n, *y = a.vecs + 7 * n;
    for (int k = tid; k < 2 * n; k += VI_THREADS) a.vecs[k] = 0.0;            // every parameter block starts at zero
    if (tid == 0) {
        st.iter = 1;                                                          // as k_ctrl_init
        st.status = 0;
        st.cur = 0;
        st.step_valid = 0;
        st.have_cand = 0;
        st.num_consecutive_invalid = 0;
        st.atleast_one_successful = 0;
        st.n_ok = st.n_bad = 0;
        st.scaling_done = 0;
        st.need_grad_check = 1;
        st.radius = a.opt.initial_radius;
        st.decrease_factor = 2.0;
        st.x_cost = st.cand_cost = st.model_cost_change = st.x_norm2 = st.initial_cost = 0.0;
        for (int k = 0; k < 64; k++) {
            st.trace_cost[k] = st.trace_radius[k] = st.trace_model[k] = 0.0;
            st.trace_accepted[k] = 0;
        }
        st.trace_radius[0] = a.opt.initial_radius;
        st.trace_accepted[0] = 1;
        acc.cost[0] = acc.cost[1] = 0.0;
        acc.model_gd = acc.model_dd = acc.step_norm2 = acc.cand_norm2 = acc.fixed_cost = 0.0;
        acc.grad_max_bits = 0ull;
        acc.schur_fail = acc.chol_fail = 0;
        if (a.max_iter <= 0) st.status = 1 + 0;
    }
    __syncthreads();
    bool first = true;
    while (true) {
        // ---- linearise at x[cur] when it changed (iteration 0, accepted steps)
        if (st.need_grad_check) {
            const double *x = xb[st.cur];
            double c = 0.0;
            for (int p = tid; p < P; p += VI_THREADS) c += viinit_eval(a, p, x, a.rw + 9 * (size_t)p, a.Jw + 81 * (size_t)p);
            c = 0.5 * viinit_block_sum(c, red);
            for (int k = tid; k < n * n; k += VI_THREADS) A[k] = 0.0;
            for (int k = tid; k < n; k += VI_THREADS) g[k] = 0.0;
            __syncthreads();
            // normal equations, pairs in order (every entry sees its contributions in pair order: deterministic)
            for (int p = 0; p < P; p++) {
                if (tid < 81) {
                    const int ca = viinit_col(a, p, tid / 9), cb = viinit_col(a, p, tid % 9);
                    if (ca >= 0 && cb >= 0) {
                        const double *Jw = a.Jw + 81 * (size_t)p;
                        double s = 0.0;
                        for (int q = 0; q < 9; q++) s += Jw[q * 9 + tid / 9] * Jw[q * 9 + tid % 9];
                        A[ca * n + cb] += s;
                    }
                } else if (tid >= 96 && tid < 105) {
                    const int ca = viinit_col(a, p, tid - 96);
                    if (ca >= 0) {
                        const double *Jw = a.Jw + 81 * (size_t)p, *rw = a.rw + 9 * (size_t)p;
                        double s = 0.0;
                        for (int q = 0; q < 9; q++) s += Jw[q * 9 + tid - 96] * rw[q];
                        g[ca] += s;
                    }
                }
                __syncthreads();
            }
            double gm = 0.0;
            for (int k = tid; k < n; k += VI_THREADS) {
                hd[k] = A[k * n + k];
                if (!st.scaling_done) scale[k] = a.opt.jacobi_scaling ? 1.0 / (1.0 + sqrt(hd[k])) : 1.0; // once, at iteration 0
                gm = fmax(gm, fabs(g[k]));
            }
            // block max through the sum helper's buffer
            for (int o = 16; o > 0; o >>= 1) gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, o));
            __syncthreads();
            if ((tid & 31) == 0) red[tid >> 5] = gm;
            __syncthreads();
            if (tid == 0) {
                for (int k = 1; k < VI_THREADS / 32; k++) gm = fmax(gm, red[k]);
                st.scaling_done = 1;
                st.need_grad_check = 0;
                if (first) {
                    st.x_cost = st.initial_cost = c;
                    st.trace_cost[0] = c;
                    acc.cost[0] = c;
                }
                if (st.status == 0 && gm <= a.opt.gradient_tolerance) st.status = 1 + 2;
            }
            first = false;
            __syncthreads();
        }
        if (st.status != 0) break;
        const int cur = st.cur;
        const double *x = xb[cur];
        double *xc = xb[1 - cur];
        // ---- damped system (H + diag(d)) delta = -g in unscaled variables, d_i = clamp(s_i^2 H_ii) / (radius s_i^2)
        for (int k = tid; k < n; k += VI_THREADS) {
            const double s2 = scale[k] * scale[k];
            damp[k] = fmin(fmax(s2 * hd[k], a.opt.min_diag), a.opt.max_diag) / (st.radius * s2);
            A[k * n + k] = hd[k] + damp[k];
            y[k] = -g[k];
        }
        if (tid == 0) fail_s = 0;
        __syncthreads();
        // right-looking Cholesky of the lower triangle; the strictly-lower part of A still holds H (the factor goes to the upper
        // triangle as L^T so that H survives a rejected step)
        for (int k = 0; k < n; k++) {
            double d = A[k * n + k];
            for (int q = 0; q < k; q++) d -= A[q * n + k] * A[q * n + k]; // L_kq = U(q,k)
            if (!(d > 0.0)) {
                if (tid == 0) fail_s = 1;
                break;
            }
            const double dk = sqrt(d), inv = 1.0 / dk;
            for (int i = k + 1 + tid; i < n; i += VI_THREADS) {
                double s = A[i * n + k];                                      // H_ik (lower triangle, untouched)
                for (int q = 0; q < k; q++) s -= A[q * n + i] * A[q * n + k];
                A[k * n + i] = s * inv;                                       // L_ik stored at U(k,i)
            }
            __syncthreads();
            if (tid == 0) A[k * n + k] = dk;                                  // diagonal of L (H_kk is kept in hd)
            __syncthreads();
        }
        __syncthreads();
        bool ok = fail_s == 0;
        if (ok) {
            // forward L z = y, backward L^T delta = z: one warp, lanes over the row
            if (tid < 32) {
                for (int k = 0; k < n; k++) {
                    double s = 0.0;
                    for (int q = tid; q < k; q += 32) s += A[q * n + k] * delta[q];
                    s = warp_sum(s);
                    if (tid == 0) delta[k] = (y[k] - s) / A[k * n + k];
                    __syncwarp();
                }
                for (int k = n - 1; k >= 0; k--) {
                    double s = 0.0;
                    for (int q = k + 1 + tid; q < n; q += 32) s += A[k * n + q] * delta[q];
                    s = warp_sum(s);
                    if (tid == 0) delta[k] = (delta[k] - s) / A[k * n + k];
                    __syncwarp();
                }
            }
            __syncthreads();
            double gd = 0.0, dd = 0.0, sn = 0.0, cn = 0.0;
            for (int k = tid; k < n; k += VI_THREADS) {
                const double dl = delta[k];
                if (!isfinite(dl)) ok = false;
                xc[k] = x[k] + dl;
                gd += g[k] * dl;
                dd += damp[k] * dl * dl;
                sn += dl * dl;
                cn += xc[k] * xc[k];
            }
            gd = viinit_block_sum(gd, red);
            dd = viinit_block_sum(dd, red);
            sn = viinit_block_sum(sn, red);
            cn = viinit_block_sum(cn, red);
            ok = __syncthreads_and(ok ? 1 : 0) != 0;
            // candidate cost, residuals only
            double c = 0.0;
            double r9[9];
            if (ok)
                for (int p = tid; p < P; p += VI_THREADS) c += viinit_eval(a, p, xc, r9, nullptr);
            c = 0.5 * viinit_block_sum(c, red);
            if (tid == 0) {
                acc.model_gd = gd;
                acc.model_dd = dd;
                acc.step_norm2 = sn;
                acc.cand_norm2 = cn;
                acc.cost[1 - cur] = c;
            }
        }
        __syncthreads();
        if (tid == 0) {
            st.step_valid = ok ? 1 : 0;
            ctrl_step(&st, &acc, a.opt, a.max_iter);
        }
        __syncthreads();
        if (st.status != 0) break;
    }
    // ---- parameter blocks out: the last accepted x
    const double *x = xb[st.cur];
    for (int k = tid; k < 3 * a.F + 3; k += VI_THREADS) a.out[k] = k < n ? x[k] : 0.0;
    {
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(&st);
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(a.st);
        for (int k = tid; k < (int)(sizeof(LMState) / 8); k += VI_THREADS) dst[k] = src[k];
        src = reinterpret_cast<const unsigned long long *>(&acc);
        dst = reinterpret_cast<unsigned long long *>(a.acc);
        for (int k = tid; k < (int)(sizeof(Accum) / 8); k += VI_THREADS) dst[k] = src[k];
    }
}

} // namespace sdv

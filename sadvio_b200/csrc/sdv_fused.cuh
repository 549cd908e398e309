// Fused visual linearisation + landmark Schur complement, and fused landmark back-substitution + candidate cost.
//
// Round 1 materialised the Jacobians (k_lin_visual wrote 160 B per observation, k_schur and k_backsub each read them back)
// and reduced into S with ~120 FP64 atomics per landmark.  Here the Jacobian of an observation lives only in the registers of
// the thread that accumulates it (a5 AngularErrCeres_pointxd_dx / a6 ReprojectionErrCeres_pointxd_dx are re-evaluated from the
// 32 B of the observation and the frame-camera table), and the contributions of a TILE of consecutive landmarks are summed in
// shared memory before they reach S:
//
//   k_lin_schur     one CTA per tile (<= FT slots, <= FT_LMK landmarks), one thread per slot = (landmark, keyframe) group:
//                   phase A  J of the slot's observations -> W_f = Jp^T Jl, H_f = Jp^T Jp, Jp^T r, partial Jl^T Jl, Jl^T r
//                   phase L  one thread per landmark: V = Jl^T Jl + D_l, V^-1, Jacobi scale (iteration 0), gradient maximum;
//                            V^-1, g_l, D_l go to lmk_aux (96 B per landmark) for the back-substitution
//                   phase S  per slot: Y_f = W_f V^-1, diagonal block H_f - Y_f W_f^T, right-hand side, into shared memory
//                   phase B  consecutive landmarks seen from the same keyframes form a RUN (found here, on the device);
//                            one thread per ENTRY of the run's (6m)^2 block sums over the landmarks of the run and issues
//                            ONE atomic: ~12 atomics per landmark at C3 instead of ~120
//   k_backsub_cost  same tiling: e = sum Jl^T (Jp delta_f) recomputed, delta_l = -V^-1 (g_l + e), candidate landmark,
//                   residual-only evaluation of the slot's observations at the candidate point -> candidate cost.
//
// Per LM iteration the visual factors now move 3 x 32 B per observation + 2 x 96 B per landmark instead of 3 x 160 B + 2 x 36 B
// per observation (SURVEY.md section 8d "whole iteration, fused").
#pragma once
#include "sdv_kernels.cuh"

namespace sdv {

// tile geometry (compile-time; -DSDV_FT / -DSDV_FT_LMK / -DSDV_FUSED_CPS for experiments: smaller tiles, more CTAs per SM)
#ifndef SDV_FT
#define SDV_FT 160
#endif
#ifndef SDV_FT_LMK
#define SDV_FT_LMK 64
#endif
#ifndef SDV_FUSED_CPS
#define SDV_FUSED_CPS 2
#endif
constexpr int FT = SDV_FT;          // threads per CTA = slots per tile (upper bound)
constexpr int FT_LMK = SDV_FT_LMK;  // landmarks per tile (upper bound)
constexpr int FUSED_CPS = SDV_FUSED_CPS;                      // resident CTAs per SM of k_lin_schur (shared memory: FT_SMEM_SCHUR each)
constexpr int FUSED_CPS_BACK = FUSED_CPS > 3 ? FUSED_CPS : 3; // ... of k_backsub_cost
constexpr int FTW = (FT + 31) / 32; // warps per CTA
static_assert(FT > FT_LMK, "one thread per landmark plus one in the tile prologue");
constexpr int FT_SD = 75;       // doubles per slot in shared memory: W 18 | Y 18 | D 39 (21 block, 6 rhs, 6 diag, 6 raw gradient)
constexpr int FT_SMEM_SCHUR = (FT * FT_SD + 9 * FT + FT_LMK * 10) * (int)sizeof(double);
constexpr int LMK_AUX = 12;     // V^-1 (6) | g_l (3) | D_l (3)

// residual only (candidate cost): the residual expressions of eval_visual, nothing else
template <int KIND> SDV_DEV void eval_residual(const double *row, const double *Kc, double w, const double *p, const double *meas, double *r) {
    const double *Rsw = row, *tsw = row + 9;
    double ts[3];
    mat3_vec(Rsw, p, ts);
    ts[0] += tsw[0];
    ts[1] += tsw[1];
    ts[2] += tsw[2];
    if (KIND == 0) {
        double nrm = norm3(ts);
        double inv = 1.0 / nrm;
        double bh[3] = {ts[0] * inv, ts[1] * inv, ts[2] * inv};
        const double *b = meas;
        double b1[3], b2[3];
        double dx0 = b[0] - 1.0;
        if (sqrt(dx0 * dx0 + b[1] * b[1] + b[2] * b[2]) > 1e-5) {
            b1[0] = 0.0; b1[1] = b[2]; b1[2] = -b[1];
        } else {
            b1[0] = b[1]; b1[1] = -b[0]; b1[2] = 0.0;
        }
        double n1 = 1.0 / norm3(b1);
        b1[0] *= n1; b1[1] *= n1; b1[2] *= n1;
        cross3(b1, b, b2);
        double n2 = 1.0 / norm3(b2);
        b2[0] *= n2; b2[1] *= n2; b2[2] *= n2;
        double e[3] = {bh[0] - b[0], bh[1] - b[1], bh[2] - b[2]};
        r[0] = w * (b1[0] * e[0] + b1[1] * e[1] + b1[2] * e[2]);
        r[1] = w * (b2[0] * e[0] + b2[1] * e[1] + b2[2] * e[2]);
    } else {
        double fx = Kc[0], fy = Kc[1], cx = Kc[2], cy = Kc[3];
        double pt0 = fx * ts[0] + cx * ts[2], pt1 = fy * ts[1] + cy * ts[2], z = ts[2];
        double u = pt0 / z, v = pt1 / z;
        bool ok = !(ts[2] < 0.1) && !(u < 0 || v < 0 || u > 2 * cx || v > 2 * cy) && isfinite(u) && isfinite(v);
        r[0] = ok ? w * (u - meas[0]) : 0.0;
        r[1] = ok ? w * (v - meas[1]) : 0.0;
    }
}

// ceres::HuberLoss(a)::Evaluate + ceres::internal::Corrector for one visual residual block (Ceres 2.2 loss_function.cc /
// corrector.cc; AOptimizer.cpp:102,223 pass a = sqrt(1.345)): rho(s) = s for s <= a^2, 2 a sqrt(s) - a^2 beyond; rho'' <= 0, so the
// corrector scales the residual AND the Jacobians by sqrt(rho') and the block's cost is rho(s) / 2.
SDV_DEV double huber_rho(double a, double s) {
    const double b = a * a;
    return s > b ? 2.0 * a * sqrt(s) - b : s;
}
SDV_DEV void huber_correct(double a, double *r, double *Jp, double *Jl) {
    const double s = r[0] * r[0] + r[1] * r[1];
    if (s > a * a) {
        const double sc = sqrt(fmax(2.2250738585072014e-308, a / sqrt(s)));
        r[0] *= sc;
        r[1] *= sc;
#pragma unroll
        for (int k = 0; k < 12; k++) Jp[k] *= sc;
#pragma unroll
        for (int k = 0; k < 6; k++) Jl[k] *= sc;
    }
}

// operands of one real observation (plane index ol < Oloc)
template <int KIND> struct ObsOperands {
    double meas[3], w;
    int fc;
};
template <int KIND> SDV_DEV void load_obs(const DevProblem &P, int ol, ObsOperands<KIND> &q) {
    constexpr int MP = KIND == 0 ? 3 : 2;
    const size_t o = (size_t)P.o0 + ol;
    q.fc = __ldg(P.obs_fc + o);
    q.meas[0] = __ldg(P.obs_meas + o * MP);
    q.meas[1] = __ldg(P.obs_meas + o * MP + 1);
    q.meas[2] = KIND == 0 ? __ldg(P.obs_meas + o * MP + 2) : 0.0;
    q.w = P.obs_w ? __ldg(P.obs_w + o) : -1.0; // -1: default weight of the camera (row[30])
}

// residual + Jacobian of entry `ol` of a slot: a real observation is re-evaluated, a pseudo-observation of a
// PoseToLandmarkFactor (k_lin_p2l) is read from the planes
template <int KIND> SDV_DEV void obs_eval(const DevProblem &P, const LinBuf &B, int ol, const double *p, double *r, double *Jp, double *Jl) {
    if (ol < P.o1 - P.o0) {
        ObsOperands<KIND> q;
        load_obs<KIND>(P, ol, q);
        const double *row = B.fct + (size_t)q.fc * FCT_ROW;
        eval_visual<KIND>(row, P.K + 4 * (q.fc % P.C), q.w < 0.0 ? row[30] : q.w, p, q.meas, r, Jp, Jl);
        if (P.huber_a > 0.0) huber_correct(P.huber_a, r, Jp, Jl); // the loss wraps the VISUAL blocks only
    } else {
        const size_t OC = (size_t)P.Ocap;
        r[0] = B.r[ol];
        r[1] = B.r[OC + ol];
#pragma unroll
        for (int k = 0; k < 12; k++) Jp[k] = B.Jp[(size_t)k * OC + ol];
#pragma unroll
        for (int k = 0; k < 6; k++) Jl[k] = B.Jl[(size_t)k * OC + ol];
    }
}

// tile-local landmark of slot t: largest li with sp[li] <= t
SDV_DEV int tile_landmark_of(const int *sp, int nl, int t) {
    int lo = 0, hi = nl - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (sp[mid] <= t) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

SDV_DEV void tri_ij(int k, int &i, int &j) { // k = i (i + 1) / 2 + j, j <= i < 6
    i = k >= 15 ? 5 : (k >= 10 ? 4 : (k >= 6 ? 3 : (k >= 3 ? 2 : (k >= 1 ? 1 : 0))));
    j = k - i * (i + 1) / 2;
}

#ifdef SDV_SCHUR_PROF
// developer trace (clock64 of thread 0 of CTA 0 and of the last CTA after every phase of their FIRST tile, globaltimer at entry / exit)
__device__ double g_schur_prof[32];
#define SCHUR_TICK(q) do { if (tid == 0 && tile == (int)blockIdx.x && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) g_schur_prof[(blockIdx.x ? 16 : 0) + (q)] = (double)(clock64() - t_entry); } while (0)
#else
#define SCHUR_TICK(q) do { } while (0)
#endif
template <int KIND>
__global__ void __launch_bounds__(FT, FUSED_CPS) k_lin_schur(const DevProblem *__restrict__ Pg, LinBuf B0, LinBuf B1, LMState *st, Accum *acc, SolverOpts opt, double *Sb,
                                                     double *scale_l, double *lmk_aux) {
    const DevProblem &P = *Pg; // device-resident problem description: the launch parameters do not depend on the window (one CUDA graph serves them all)
#ifdef SDV_SCHUR_PROF
    const long long t_entry = clock64();
    unsigned long long gt0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
#endif
    if (st->status != 0) return;
    const LinBuf &B = st->cur ? B1 : B0;
    extern __shared__ double fsm[];
    double *slotd = fsm;             // [FT][FT_SD]
    double *part = fsm + FT * FT_SD; // [9][FT]: partial Jl^T Jl (6) and Jl^T r (3) of every slot
    double *lmd = part + 9 * FT;     // [FT_LMK][10]: V^-1 (6), g_l (3), eliminated flag
    __shared__ int sp[FT_LMK + 1], scol[FT], sframe[FT], rbeg[FT_LMK + 1];
    __shared__ unsigned char rflag[FT_LMK];
    __shared__ int s_nruns;
    __shared__ double gred[FTW];
    const int tid = threadIdx.x;
    const int ld = P.ld;
    double *g = Sb + (size_t)P.n_pad * ld, *cdiag = g + ld, *graw = cdiag + ld;
    const double radius = st->radius;
    const bool first = st->scaling_done == 0;
    double gmax = 0.0;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        const int lA = P.tile_ptr[tile], lB = P.tile_ptr[tile + 1], nl = lB - lA;
        const int s0 = P.slot_ptr[lA];
        if (tid <= nl) sp[tid] = P.slot_ptr[lA + tid] - s0;
        __syncthreads();
        const int ns = sp[nl];
        const bool active = tid < ns;
        SCHUR_TICK(0);
        // ---------------------------------------------------------------- phase A: one thread per slot
        SlotAcc a;
        double hl[6] = {0, 0, 0, 0, 0, 0}, gl[3] = {0, 0, 0};
        int li = 0, col = -1, dc = -1;
#pragma unroll
        for (int k = 0; k < 18; k++) a.W[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 21; k++) a.H[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 6; k++) a.gp[k] = 0.0;
        if (active) {
            li = tile_landmark_of(sp, nl, tid);
            const int l = lA + li, s = s0 + tid;
            const int f = P.slot_frame[s];
            col = P.pose_col[f];
            scol[tid] = col;
            sframe[tid] = f;
            dc = P.lmk_col[l];
            double p[3];
            landmark_position(P, B, l, p);
            const int qa = P.slot_obs_ptr[s], qb = P.slot_obs_ptr[s + 1];
            SCHUR_TICK(1);
            for (int q = qa; q < qb; q++) {
                double r[2], Jp[12], Jl[6];
                obs_eval<KIND>(P, B, P.slot_obs[q], p, r, Jp, Jl);
#pragma unroll
                for (int i = 0; i < 6; i++) {
#pragma unroll
                    for (int j = 0; j < 3; j++) a.W[i * 3 + j] += Jp[i] * Jl[j] + Jp[6 + i] * Jl[3 + j];
#pragma unroll
                    for (int j = 0; j <= i; j++) a.H[tri_idx(i, j)] += Jp[i] * Jp[j] + Jp[6 + i] * Jp[6 + j];
                    a.gp[i] += Jp[i] * r[0] + Jp[6 + i] * r[1];
                }
                hl[0] += Jl[0] * Jl[0] + Jl[3] * Jl[3];
                hl[1] += Jl[0] * Jl[1] + Jl[3] * Jl[4];
                hl[2] += Jl[0] * Jl[2] + Jl[3] * Jl[5];
                hl[3] += Jl[1] * Jl[1] + Jl[4] * Jl[4];
                hl[4] += Jl[1] * Jl[2] + Jl[4] * Jl[5];
                hl[5] += Jl[2] * Jl[2] + Jl[5] * Jl[5];
                gl[0] += Jl[0] * r[0] + Jl[3] * r[1];
                gl[1] += Jl[1] * r[0] + Jl[4] * r[1];
                gl[2] += Jl[2] * r[0] + Jl[5] * r[1];
            }
        }
#pragma unroll
        for (int k = 0; k < 6; k++) part[k * FT + tid] = hl[k];
#pragma unroll
        for (int k = 0; k < 3; k++) part[(6 + k) * FT + tid] = gl[k];
        SCHUR_TICK(2);
        __syncthreads();
        SCHUR_TICK(3);
        // ---------------------------------------------------------------- phase L: one thread per landmark
        if (tid < nl) {
            const int l = lA + tid, t0 = sp[tid], t1 = sp[tid + 1], m = t1 - t0;
            double h6[6] = {0, 0, 0, 0, 0, 0}, g3[3] = {0, 0, 0};
            for (int t = t0; t < t1; t++) {
#pragma unroll
                for (int k = 0; k < 6; k++) h6[k] += part[k * FT + t];
#pragma unroll
                for (int k = 0; k < 3; k++) g3[k] += part[(6 + k) * FT + t];
            }
            const int dcl = P.lmk_col[l];
            const bool kept = dcl >= 0;
            // a run = consecutive eliminated landmarks seen from the same keyframes in the same slot order
            bool start = tid == 0 || kept || P.lmk_col[l - 1] >= 0 || m != t0 - sp[tid - 1];
            if (!start)
                for (int q = 0; q < m; q++) start |= sframe[t0 + q] != sframe[sp[tid - 1] + q];
            rflag[tid] = start ? (kept ? 2 : 1) : 0;
            double Vi[6] = {0, 0, 0, 0, 0, 0};
            bool elim = false;
            if (P.lmk_const) {
                // SetParameterBlockConstant on every landmark (addSingleFrameResiduals, …Analytic.cpp:41-43): no columns, no
                // elimination — with V^-1 = 0 and g_l = 0 the slot's Schur products reduce to H_f and Jp^T r, and the
                // back-substitution leaves the landmark where it is
                elim = true;
#pragma unroll
                for (int k = 0; k < 3; k++) g3[k] = 0.0;
                double *ax = lmk_aux + (size_t)LMK_AUX * l;
#pragma unroll
                for (int k = 0; k < LMK_AUX; k++) ax[k] = 0.0;
            } else if (kept) {
                // kept (dense) landmark: its columns live in the reduced system, no elimination
                const int ii[6] = {0, 1, 2, 1, 2, 2}, jj[6] = {0, 0, 0, 1, 1, 2};
                for (int k = 0; k < 6; k++) atomicAdd(&Sb[(size_t)(dcl + ii[k]) * ld + dcl + jj[k]], h6[k]);
                atomicAdd(&cdiag[dcl + 0], h6[0]);
                atomicAdd(&cdiag[dcl + 1], h6[3]);
                atomicAdd(&cdiag[dcl + 2], h6[5]);
                for (int k = 0; k < 3; k++) {
                    atomicAdd(&g[dcl + k], g3[k]);
                    atomicAdd(&graw[dcl + k], g3[k]);
                }
            } else {
                double s3[3];
                if (first) {
                    s3[0] = opt.jacobi_scaling ? 1.0 / (1.0 + sqrt(h6[0])) : 1.0;
                    s3[1] = opt.jacobi_scaling ? 1.0 / (1.0 + sqrt(h6[3])) : 1.0;
                    s3[2] = opt.jacobi_scaling ? 1.0 / (1.0 + sqrt(h6[5])) : 1.0;
                    scale_l[3 * (size_t)l] = s3[0];
                    scale_l[3 * (size_t)l + 1] = s3[1];
                    scale_l[3 * (size_t)l + 2] = s3[2];
                } else {
                    s3[0] = scale_l[3 * (size_t)l];
                    s3[1] = scale_l[3 * (size_t)l + 1];
                    s3[2] = scale_l[3 * (size_t)l + 2];
                }
                gmax = fmax(gmax, fmax(fabs(g3[0]), fmax(fabs(g3[1]), fabs(g3[2]))));
                const double d3[3] = {lm_damping(h6[0], s3[0], radius, opt), lm_damping(h6[3], s3[1], radius, opt), lm_damping(h6[5], s3[2], radius, opt)};
                const double V[6] = {h6[0] + d3[0], h6[1], h6[2], h6[3] + d3[1], h6[4], h6[5] + d3[2]};
                elim = sym3_inverse(V, Vi);
                if (!elim) acc->schur_fail = 1;
                double *ax = lmk_aux + (size_t)LMK_AUX * l;
#pragma unroll
                for (int k = 0; k < 6; k++) ax[k] = Vi[k];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    ax[6 + k] = g3[k];
                    ax[9 + k] = d3[k];
                }
            }
#pragma unroll
            for (int k = 0; k < 6; k++) lmd[tid * 10 + k] = Vi[k];
#pragma unroll
            for (int k = 0; k < 3; k++) lmd[tid * 10 + 6 + k] = g3[k];
            lmd[tid * 10 + 9] = elim ? 1.0 : 0.0;
        }
        SCHUR_TICK(4);
        __syncthreads();
        SCHUR_TICK(5);
        // ---------------------------------------------------------------- phase S: per slot, Schur products into shared memory
        if (active) {
            double *sd = slotd + tid * FT_SD;
            const double *lm = lmd + li * 10;
            if (dc >= 0) {
                if (col >= 0) {
#pragma unroll
                    for (int i = 0; i < 6; i++) {
#pragma unroll
                        for (int j = 0; j <= i; j++) atomicAdd(&Sb[(size_t)(col + i) * ld + col + j], a.H[tri_idx(i, j)]);
                        atomicAdd(&g[col + i], a.gp[i]);
                        atomicAdd(&cdiag[col + i], a.H[tri_idx(i, i)]);
                        atomicAdd(&graw[col + i], a.gp[i]);
#pragma unroll
                        for (int j = 0; j < 3; j++) { // dense landmark columns come after every frame column
                            if (dc > col) atomicAdd(&Sb[(size_t)(dc + j) * ld + col + i], a.W[i * 3 + j]);
                            else atomicAdd(&Sb[(size_t)(col + i) * ld + dc + j], a.W[i * 3 + j]);
                        }
                    }
                }
            } else if (lm[9] != 0.0) {
                const double v0 = lm[0], v1 = lm[1], v2 = lm[2], v3 = lm[3], v4 = lm[4], v5 = lm[5];
                const double g0 = lm[6], g1 = lm[7], g2 = lm[8];
                double Y[18];
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    const double w0 = a.W[i * 3], w1 = a.W[i * 3 + 1], w2 = a.W[i * 3 + 2];
                    Y[i * 3] = w0 * v0 + w1 * v1 + w2 * v2;
                    Y[i * 3 + 1] = w0 * v1 + w1 * v3 + w2 * v4;
                    Y[i * 3 + 2] = w0 * v2 + w1 * v4 + w2 * v5;
                }
#pragma unroll
                for (int k = 0; k < 18; k++) {
                    sd[k] = a.W[k];
                    sd[18 + k] = Y[k];
                }
#pragma unroll
                for (int i = 0; i < 6; i++) {
#pragma unroll
                    for (int j = 0; j <= i; j++)
                        sd[36 + tri_idx(i, j)] = a.H[tri_idx(i, j)] - (Y[i * 3] * a.W[j * 3] + Y[i * 3 + 1] * a.W[j * 3 + 1] + Y[i * 3 + 2] * a.W[j * 3 + 2]);
                    sd[57 + i] = a.gp[i] - (Y[i * 3] * g0 + Y[i * 3 + 1] * g1 + Y[i * 3 + 2] * g2);
                    sd[63 + i] = a.H[tri_idx(i, i)];
                    sd[69 + i] = a.gp[i];
                }
            } else {
#pragma unroll
                for (int k = 0; k < FT_SD; k++) sd[k] = 0.0; // 3x3 block not invertible: the step is invalid anyway (Accum::schur_fail)
            }
        }
        if (tid == 0) {
            int nr = 0;
            for (int q = 0; q < nl; q++)
                if (rflag[q]) rbeg[nr++] = q;
            rbeg[nr] = nl;
            s_nruns = nr;
        }
        SCHUR_TICK(6);
        __syncthreads();
        SCHUR_TICK(7);
        // ---------------------------------------------------------------- phase B: one thread per entry of a run's block
        const int nruns = s_nruns;
        for (int r = 0; r < nruns; r++) {
            const int la = rbeg[r], lb = rbeg[r + 1];
            if (rflag[la] == 2) continue; // kept landmark: went straight to S
            const int t0 = sp[la], m = sp[la + 1] - t0;
            const int ediag = 39 * m, etot = P.lmk_const ? ediag : ediag + 18 * m * (m - 1); // constant landmarks couple nothing
            // every landmark of a run has m slots: landmark q of the run starts m * FT_SD doubles after the previous one
            const int nrun = lb - la, lstride = m * FT_SD;
            const double *base0 = slotd + t0 * FT_SD;
            for (int e = tid; e < etot; e += FT) {
                double s = 0.0;
                double *dst = nullptr;
                if (e < ediag) {
                    const int sa = e / 39, k = e - sa * 39, ca = scol[t0 + sa];
                    if (ca < 0) continue;
                    const double *src = base0 + sa * FT_SD + 36 + k;
                    for (int q = 0; q < nrun; q++, src += lstride) s += *src;
                    if (k < 21) {
                        int i, j;
                        tri_ij(k, i, j);
                        dst = &Sb[(size_t)(ca + i) * ld + ca + j];
                    } else if (k < 27) dst = &g[ca + k - 21];
                    else if (k < 33) dst = &cdiag[ca + k - 27];
                    else dst = &graw[ca + k - 33];
                } else {
                    const int e2 = e - ediag, pi = e2 / 36, ij = e2 - pi * 36, i = ij / 6, j = ij - i * 6;
                    int sa = 1, base = 0;
                    while (base + sa <= pi) {
                        base += sa;
                        sa++;
                    }
                    const int sb = pi - base;
                    const int ca = scol[t0 + sa], cb = scol[t0 + sb];
                    if (ca < 0 || cb < 0) continue;
                    const double *Ya = base0 + sa * FT_SD + 18 + i * 3, *Wb = base0 + sb * FT_SD + j * 3;
                    double s1 = 0.0, s2 = 0.0;
                    for (int q = 0; q < nrun; q++, Ya += lstride, Wb += lstride) {
                        s = fma(Ya[0], Wb[0], s);
                        s1 = fma(Ya[1], Wb[1], s1);
                        s2 = fma(Ya[2], Wb[2], s2);
                    }
                    s = -(s + s1 + s2);
                    dst = ca > cb ? &Sb[(size_t)(ca + i) * ld + cb + j] : &Sb[(size_t)(cb + j) * ld + ca + i];
                }
                atomicAdd(dst, s);
            }
        }
        SCHUR_TICK(8);
        __syncthreads(); // the next tile reuses every shared array
        SCHUR_TICK(9);
    }
#ifdef SDV_SCHUR_PROF
    if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {
        unsigned long long gt1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
        g_schur_prof[(blockIdx.x ? 16 : 0) + 10] = (double)(clock64() - t_entry);
        g_schur_prof[(blockIdx.x ? 16 : 0) + 11] = (double)gt0;
        g_schur_prof[(blockIdx.x ? 16 : 0) + 12] = (double)gt1;
    }
#endif
    for (int o = 16; o > 0; o >>= 1) gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    if ((tid & 31) == 0) gred[tid >> 5] = gmax;
    __syncthreads();
    if (tid == 0) {
        for (int q = 1; q < FTW; q++) gmax = fmax(gmax, gred[q]);
        if (gmax > 0.0) atomic_max_nonneg(reinterpret_cast<double *>(&acc->grad_max_bits), gmax);
    }
}

// landmark back-substitution delta_l = -V^-1 (g_l + sum_f W_f^T delta_f) with W_f^T delta_f = sum_obs Jl^T (Jp delta_f) recomputed,
// candidate landmark parameters, model-decrease / norm partial sums, and the cost of the visual factors at the candidate point
template <int KIND>
// Side job (Sz / nz): the reduced system S of this iteration has been consumed by the factorisation that ran before this kernel and
// the next iteration accumulates into it from zero — the grid clears it here (a few 16-byte stores per thread) instead of a
// memset node of its own at the head of every iteration.
__global__ void __launch_bounds__(FT, FUSED_CPS_BACK) k_backsub_cost(const DevProblem *__restrict__ Pg, LinBuf B0, LinBuf B1, const LMState *st, Accum *acc, const double *dxp,
                                                        const double *lmk_aux, double *Sz = nullptr, long long nz = 0) {
    const DevProblem &P = *Pg; // device-resident problem description: the launch parameters do not depend on the window (one CUDA graph serves them all)
    if (st->status != 0) return;
    for (long long i = 2 * ((long long)blockIdx.x * FT + threadIdx.x); i < nz; i += 2 * (long long)gridDim.x * FT)
        *reinterpret_cast<double2 *>(Sz + i) = make_double2(0.0, 0.0);
    if (!st->step_valid) return;
    const int cand = 1 - st->cur;
    const LinBuf &Bx = st->cur ? B1 : B0;
    const LinBuf &Bc = st->cur ? B0 : B1;
    __shared__ double part[3][FT], pc[FT_LMK][3], red[FTW][5];
    __shared__ int sp[FT_LMK + 1];
    const int tid = threadIdx.x;
    const int Oloc = P.o1 - P.o0;
    double gd = 0, dd = 0, sn = 0, cn = 0, cost = 0;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        const int lA = P.tile_ptr[tile], lB = P.tile_ptr[tile + 1], nl = lB - lA;
        const int s0 = P.slot_ptr[lA];
        if (tid <= nl) sp[tid] = P.slot_ptr[lA + tid] - s0;
        __syncthreads();
        const int ns = sp[nl];
        const bool active = tid < ns;
        int li = 0, qa = 0, qb = 0;
        double e[3] = {0, 0, 0};
        if (active) {
            li = tile_landmark_of(sp, nl, tid);
            const int l = lA + li, s = s0 + tid;
            qa = P.slot_obs_ptr[s];
            qb = P.slot_obs_ptr[s + 1];
            if (P.lmk_col[l] < 0 && !P.lmk_const) {
                const int pcol = P.pose_col[P.slot_frame[s]];
                if (pcol >= 0) { // a constant keyframe does not move: no contribution
                    double d[6], p[3];
#pragma unroll
                    for (int k = 0; k < 6; k++) d[k] = dxp[pcol + k];
                    landmark_position(P, Bx, l, p);
                    for (int q = qa; q < qb; q++) {
                        double r[2], Jp[12], Jl[6];
                        obs_eval<KIND>(P, Bx, P.slot_obs[q], p, r, Jp, Jl);
                        double u0 = 0, u1 = 0;
#pragma unroll
                        for (int k = 0; k < 6; k++) {
                            u0 += Jp[k] * d[k];
                            u1 += Jp[6 + k] * d[k];
                        }
                        e[0] += Jl[0] * u0 + Jl[3] * u1;
                        e[1] += Jl[1] * u0 + Jl[4] * u1;
                        e[2] += Jl[2] * u0 + Jl[5] * u1;
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 3; k++) part[k][tid] = e[k];
        __syncthreads();
        if (tid < nl) {
            const int l = lA + tid;
            const int dc = P.lmk_col[l];
            double xc[3];
            if (dc >= 0) {
                // kept landmark: part of the reduced system; mirror it so that every landmark is read from xl
#pragma unroll
                for (int k = 0; k < 3; k++) xc[k] = Bc.xp[dc + k];
            } else {
                double es[3] = {0, 0, 0};
                for (int t = sp[tid]; t < sp[tid + 1]; t++) {
                    es[0] += part[0][t];
                    es[1] += part[1][t];
                    es[2] += part[2][t];
                }
                const double *ax = lmk_aux + (size_t)LMK_AUX * l;
                const double t3[3] = {ax[6] + es[0], ax[7] + es[1], ax[8] + es[2]};
                const double dl[3] = {-(ax[0] * t3[0] + ax[1] * t3[1] + ax[2] * t3[2]), -(ax[1] * t3[0] + ax[3] * t3[1] + ax[4] * t3[2]),
                                      -(ax[2] * t3[0] + ax[4] * t3[1] + ax[5] * t3[2])};
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    xc[k] = Bx.xl[3 * (size_t)l + k] + dl[k];
                    gd += ax[6 + k] * dl[k];
                    dd += ax[9 + k] * dl[k] * dl[k];
                    sn += dl[k] * dl[k];
                    cn += xc[k] * xc[k];
                }
            }
#pragma unroll
            for (int k = 0; k < 3; k++) {
                Bc.xl[3 * (size_t)l + k] = xc[k];
                pc[tid][k] = P.lmk_t[3 * (size_t)l + k] + xc[k];
            }
        }
        __syncthreads();
        if (active) {
            const double p[3] = {pc[li][0], pc[li][1], pc[li][2]};
            for (int q = qa; q < qb; q++) {
                const int ol = P.slot_obs[q];
                if (ol >= Oloc) continue; // pseudo-observations: k_lin_p2l adds their candidate cost
                ObsOperands<KIND> o;
                load_obs<KIND>(P, ol, o);
                const double *row = Bc.fct + (size_t)o.fc * FCT_ROW;
                if (P.lmk_const && P.pose_col[o.fc / P.C] < 0) continue; // every block constant: part of fixed_cost, not of the cost
                double r[2];
                eval_residual<KIND>(row, P.K + 4 * (o.fc % P.C), o.w < 0.0 ? row[30] : o.w, p, o.meas, r);
                const double s2 = r[0] * r[0] + r[1] * r[1];
                cost += P.huber_a > 0.0 ? huber_rho(P.huber_a, s2) : s2;
            }
        }
        __syncthreads();
    }
    gd = warp_sum(gd);
    dd = warp_sum(dd);
    sn = warp_sum(sn);
    cn = warp_sum(cn);
    cost = warp_sum(cost);
    if ((tid & 31) == 0) {
        red[tid >> 5][0] = gd;
        red[tid >> 5][1] = dd;
        red[tid >> 5][2] = sn;
        red[tid >> 5][3] = cn;
        red[tid >> 5][4] = cost;
    }
    __syncthreads();
    if (tid < 5) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < FTW; w++) v += red[w][tid];
        if (tid == 4) v *= 0.5;
        double *dst = tid == 0 ? &acc->model_gd : (tid == 1 ? &acc->model_dd : (tid == 2 ? &acc->step_norm2 : (tid == 3 ? &acc->cand_norm2 : &acc->cost[cand])));
        if (v != 0.0) atomicAdd(dst, v);
    }
}

// cost of the visual factors at the point of one linearisation buffer (iteration 0): residual-only sweep, one thread per
// observation, nothing written but Accum::cost
template <int KIND> __global__ void __launch_bounds__(256) k_visual_cost(const DevProblem *__restrict__ Pg, LinBuf B0, LinBuf B1, const LMState *st, Accum *acc, int which) {
    const DevProblem &P = *Pg; // device-resident problem description: the launch parameters do not depend on the window (one CUDA graph serves them all)
    if (st->status != 0) return;
    const int b = which >= 0 ? which : (which == -1 ? st->cur : 1 - st->cur);
    const LinBuf &B = b ? B1 : B0;
    __shared__ double red[8], redf[8];
    const int Oloc = P.o1 - P.o0;
    double cost = 0.0, fixed = 0.0;
    for (int ol = blockIdx.x * blockDim.x + threadIdx.x; ol < Oloc; ol += gridDim.x * blockDim.x) {
        ObsOperands<KIND> o;
        load_obs<KIND>(P, ol, o);
        const int l = __ldg(P.obs_lmk + P.o0 + ol);
        double p[3];
        landmark_position(P, B, l, p);
        const double *row = B.fct + (size_t)o.fc * FCT_ROW;
        double r[2];
        eval_residual<KIND>(row, P.K + 4 * (o.fc % P.C), o.w < 0.0 ? row[30] : o.w, p, o.meas, r);
        double s2 = r[0] * r[0] + r[1] * r[1];
        if (P.huber_a > 0.0) s2 = huber_rho(P.huber_a, s2);
        // constant landmark seen from a constant keyframe: the residual block has no free parameter block (Ceres fixed_cost)
        if (P.lmk_const && P.pose_col[o.fc / P.C] < 0) fixed += s2;
        else cost += s2;
    }
    cost = warp_sum(cost);
    fixed = warp_sum(fixed);
    if ((threadIdx.x & 31) == 0) {
        red[threadIdx.x >> 5] = cost;
        redf[threadIdx.x >> 5] = fixed;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0, sf = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) {
            s += red[i];
            sf += redf[i];
        }
        if (s != 0.0) atomicAdd(&acc->cost[b], 0.5 * s);
        if (sf != 0.0 && which == 0) atomicAdd(&acc->fixed_cost, 0.5 * sf); // accumulated during the first linearisation only
    }
}

} // namespace sdv

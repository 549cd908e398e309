// sm_100a kernels of the sliding-window BA/VIO solve.  One LM iteration =
//   k_lin_schur (sdv_fused.cuh) + k_assemble_factors -> (NCCL all-reduce of the reduced system) -> k_chol_band (or k_sysprep +
//   k_chol_chain / k_chol_panel x nT + k_trisolve) -> k_backsub_cost + k_lin_factors (candidate) -> k_ctrl
// Every kernel reads LMState::status first and returns when the solve has terminated, so the fixed launch sequence
// can be replayed (and graph-captured) without host round trips.
#pragma once
#include "sdv_math.cuh"
#include "sdv_types.cuh"

namespace sdv {

// =====================================================================================================================
// frame-camera table
// =====================================================================================================================
SDV_DEV void compute_fct_row(const DevProblem &P, const double *xp, int f, int c, double *row) {
    double dx[6] = {0, 0, 0, 0, 0, 0};
    int pc = P.pose_col[f];
    if (pc >= 0) {
#pragma unroll
        for (int k = 0; k < 6; k++) dx[k] = xp[pc + k];
    }
    const double *T = P.T_f_w + 12 * f;
    const double *S = P.T_s_f + 12 * c;
    double Rb[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
    double tb[3] = {T[3], T[7], T[11]};
    double Rs[9] = {S[0], S[1], S[2], S[4], S[5], S[6], S[8], S[9], S[10]};
    double ts[3] = {S[3], S[7], S[11]};
    double dR[9], Rfw[9], tfw[3], tmp[3];
    exp_so3(dx, dR);
    mat3_mul(Rb, dR, Rfw);
    mat3_vec(Rb, dx + 3, tmp);
    tfw[0] = tmp[0] + tb[0];
    tfw[1] = tmp[1] + tb[1];
    tfw[2] = tmp[2] + tb[2];
    double Rsw[9], tsw[3], G[9], Jq[9];
    mat3_mul(Rs, Rfw, Rsw);
    mat3_vec(Rs, tfw, tsw);
    tsw[0] += ts[0];
    tsw[1] += ts[1];
    tsw[2] += ts[2];
    mat3_mul(Rs, Rb, G);
    if (P.kind == 0) {
        // AngularAdjustmentCERESAnalytic.h:97-98: so3_rightJacobian(se3_RTtoVec6d(dT).block<3,1>(0,0))
        double w2[3];
        log_so3(dR, w2);
        right_jacobian(w2, Jq);
    } else {
        // Camera.cpp:109-110 followed by BundleAdjustmentCERESAnalytic.h:73-78:
        //   Jr(log R') * ( Jr(log R')^-1 * Jr(dw) ),  R' = R_f_w exp(dw)
        double wl[3], A[9], Ai[9], B[9], AB[9];
        log_so3(Rfw, wl);
        right_jacobian(wl, A);
        inverse3(A, Ai);
        right_jacobian(dx, B);
        mat3_mul(Ai, B, AB);
        mat3_mul(A, AB, Jq);
    }
#pragma unroll
    for (int k = 0; k < 9; k++) {
        row[k] = Rsw[k];
        row[12 + k] = G[k];
        row[21 + k] = Jq[k];
    }
    row[9] = tsw[0];
    row[10] = tsw[1];
    row[11] = tsw[2];
    row[30] = P.cam_w[c];
    row[31] = 0.0;
    row[32] = 0.0;
    row[33] = 0.0;
}

__global__ void k_prep_table(const DevProblem *__restrict__ Pg, LinBuf B0, LinBuf B1, const LMState *st, int which /* -1: cur, -2: 1-cur, else fixed */) {
    const DevProblem &P = *Pg; // device-resident problem description: the launch parameters do not depend on the window (one CUDA graph serves them all)
    if (st->status != 0) return;
    int b = which >= 0 ? which : (which == -1 ? st->cur : 1 - st->cur);
    LinBuf &B = b ? B1 : B0;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.F * P.C) return;
    compute_fct_row(P, B.xp, i / P.C, i % P.C, B.fct + (size_t)i * FCT_ROW);
}

// =====================================================================================================================
// visual residual + Jacobian  (a5: AngularErrCeres_pointxd_dx, a6: ReprojectionErrCeres_pointxd_dx + Camera::project)
// =====================================================================================================================
template <int KIND>
SDV_DEV void eval_visual(const double *row, const double *Kc, double w, const double *p, const double *meas, double *r,
                         double *Jp, double *Jl) {
    const double *Rsw = row, *tsw = row + 9, *G = row + 12, *Jq = row + 21;
    double ts[3];
    mat3_vec(Rsw, p, ts);
    ts[0] += tsw[0];
    ts[1] += tsw[1];
    ts[2] += tsw[2];
    double A[6]; // 2x3: d r / d t_s
    if (KIND == 0) {
        double nrm = norm3(ts);
        double inv = 1.0 / nrm;
        double bh[3] = {ts[0] * inv, ts[1] * inv, ts[2] * inv};
        const double *b = meas;
        // tangent basis (AngularAdjustmentCERESAnalytic.h:67-77)
        double b1[3], b2[3];
        double dx0 = b[0] - 1.0;
        if (sqrt(dx0 * dx0 + b[1] * b[1] + b[2] * b[2]) > 1e-5) {
            b1[0] = 0.0; b1[1] = b[2]; b1[2] = -b[1];       // b x (1,0,0)
        } else {
            b1[0] = b[1]; b1[1] = -b[0]; b1[2] = 0.0;       // b x (0,0,1)
        }
        double n1 = 1.0 / norm3(b1);
        b1[0] *= n1; b1[1] *= n1; b1[2] *= n1;
        cross3(b1, b, b2);
        double n2 = 1.0 / norm3(b2);
        b2[0] *= n2; b2[1] *= n2; b2[2] *= n2;
        double e[3] = {bh[0] - b[0], bh[1] - b[1], bh[2] - b[2]};
        r[0] = w * (b1[0] * e[0] + b1[1] * e[1] + b1[2] * e[2]);
        r[1] = w * (b2[0] * e[0] + b2[1] * e[1] + b2[2] * e[2]);
        double d1 = b1[0] * bh[0] + b1[1] * bh[1] + b1[2] * bh[2];
        double d2 = b2[0] * bh[0] + b2[1] * bh[1] + b2[2] * bh[2];
        double s = w * inv;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            A[k] = s * (b1[k] - d1 * bh[k]);
            A[3 + k] = s * (b2[k] - d2 * bh[k]);
        }
    } else {
        // Camera::project, Camera.cpp:92-101,128-136
        double fx = Kc[0], fy = Kc[1], cx = Kc[2], cy = Kc[3];
        double pt0 = fx * ts[0] + cx * ts[2], pt1 = fy * ts[1] + cy * ts[2], z = ts[2];
        double iz = 1.0 / z;
        double u = pt0 / z, v = pt1 / z;
        bool ok = !(ts[2] < 0.1) && !(u < 0 || v < 0 || u > 2 * cx || v > 2 * cy) && isfinite(u) && isfinite(v);
        r[0] = ok ? w * (u - meas[0]) : 0.0; // failed projection: residual zeroed, Jacobian kept (BA…Analytic.h:63-68)
        r[1] = ok ? w * (v - meas[1]) : 0.0;
        double jh02 = -pt0 / (z * z), jh12 = -pt1 / (z * z);
        A[0] = w * (iz * fx); A[1] = 0.0;          A[2] = w * (iz * cx + jh02);
        A[3] = 0.0;           A[4] = w * (iz * fy); A[5] = w * (iz * cy + jh12);
    }
    // J_lmk = A * Rsw ; J_trans = A * G ; J_rot = -(J_lmk [p]x) Jq
#pragma unroll
    for (int i = 0; i < 2; i++) {
        double a0 = A[i * 3], a1 = A[i * 3 + 1], a2 = A[i * 3 + 2];
        double l0 = a0 * Rsw[0] + a1 * Rsw[3] + a2 * Rsw[6];
        double l1 = a0 * Rsw[1] + a1 * Rsw[4] + a2 * Rsw[7];
        double l2 = a0 * Rsw[2] + a1 * Rsw[5] + a2 * Rsw[8];
        Jl[i * 3] = l0; Jl[i * 3 + 1] = l1; Jl[i * 3 + 2] = l2;
        Jp[i * 6 + 3] = a0 * G[0] + a1 * G[3] + a2 * G[6];
        Jp[i * 6 + 4] = a0 * G[1] + a1 * G[4] + a2 * G[7];
        Jp[i * 6 + 5] = a0 * G[2] + a1 * G[5] + a2 * G[8];
        // row-vector l * [p]x = l x p
        double u0 = l1 * p[2] - l2 * p[1], u1 = l2 * p[0] - l0 * p[2], u2 = l0 * p[1] - l1 * p[0];
        Jp[i * 6 + 0] = -(u0 * Jq[0] + u1 * Jq[3] + u2 * Jq[6]);
        Jp[i * 6 + 1] = -(u0 * Jq[1] + u1 * Jq[4] + u2 * Jq[7]);
        Jp[i * 6 + 2] = -(u0 * Jq[2] + u1 * Jq[5] + u2 * Jq[8]);
    }
}

SDV_DEV void landmark_position(const DevProblem &P, const LinBuf &B, int l, double *p) {
    int dc = P.lmk_col[l];
    const double *d = dc >= 0 ? B.xp + dc : B.xl + 3 * (size_t)l;
    p[0] = P.lmk_t[3 * (size_t)l] + d[0];
    p[1] = P.lmk_t[3 * (size_t)l + 1] + d[1];
    p[2] = P.lmk_t[3 * (size_t)l + 2] + d[2];
}

// mbarrier / bulk-copy (TMA engine, 1-D) helpers ------------------------------------------------------------------
SDV_DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
SDV_DEV void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
SDV_DEV void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
SDV_DEV void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
SDV_DEV void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

constexpr int LIN_THREADS = 256;

// Evaluate every visual residual block of this rank at the point held by the chosen linearisation buffer, write
// r / J_pose / J_lmk as SoA planes, and add 1/2 sum r^2 to Accum::cost[buf].  Persistent grid: each CTA stages the
// frame-camera table into shared memory once with a TMA bulk copy and then walks observation tiles.
template <int KIND, bool SMEM, bool EARLY>
__global__ void __launch_bounds__(LIN_THREADS) k_lin_visual(const DevProblem *__restrict__ Pg, LinBuf B0, LinBuf B1, const LMState *st, Accum *acc,
                                                            int which) {
    const DevProblem &P = *Pg; // device-resident problem description: the launch parameters do not depend on the window (one CUDA graph serves them all)
    if (st->status != 0) return;
    if (which == -2 && !st->step_valid) return; // no candidate to evaluate after an invalid step
    int b = which >= 0 ? which : (which == -1 ? st->cur : 1 - st->cur);
    const LinBuf &B = b ? B1 : B0;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ double red[LIN_THREADS / 32];
    const double *fct = B.fct;
    const int Oloc = P.o1 - P.o0, OC = P.Ocap;
    constexpr int fstride = FCT_ROW;
    if constexpr (SMEM) {
        // stage the whole table with a few large TMA bulk copies (same padded layout in global and shared memory); the wait
        // comes after the pipeline prologue below so the first index / operand loads overlap the copy
        const uint32_t total = (uint32_t)(P.F * P.C * FCT_ROW * sizeof(double));
        constexpr uint32_t CHUNK = 16384;
        if (threadIdx.x == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) mbar_expect_tx(&bar, total);
        if (threadIdx.x < 32)
            for (uint32_t off = threadIdx.x * CHUNK; off < total; off += 32 * CHUNK)
                bulk_g2s(smem_raw + off, reinterpret_cast<const unsigned char *>(B.fct) + off, min(CHUNK, total - off), &bar);
        fct = reinterpret_cast<const double *>(smem_raw);
    }
    // Three-stage software pipeline over this thread's observations (stride = grid size): the indices of tile t+2 and the
    // gathered operands of tile t+1 are in flight while tile t is evaluated, so the two dependent memory round trips
    // (index -> landmark / measurement) overlap the FP64 work instead of serialising with it.
    constexpr int MP = KIND == 0 ? 3 : 2; // bearing[3] or uv[2], array-of-structs
    const int stride = gridDim.x * LIN_THREADS;
    double csum = 0.0;
    int ol = blockIdx.x * LIN_THREADS + threadIdx.x;
    int lA = 0, fcA = 0;                       // stage A: indices
    int lB = 0, fcB = 0;                       // stage B: gathered operands
    double mB[3] = {0, 0, 0}, tB[3] = {0, 0, 0}, xB[3] = {0, 0, 0}, wB = 0.0;
    auto load_idx = [&](int i) {
        lA = __ldg(P.obs_lmk + P.o0 + i);
        fcA = __ldg(P.obs_fc + P.o0 + i);
    };
    auto gather = [&](int i) {
        const size_t o = (size_t)P.o0 + i;
        lB = lA;
        fcB = fcA;
        mB[0] = __ldg(P.obs_meas + o * MP);
        mB[1] = __ldg(P.obs_meas + o * MP + 1);
        if (KIND == 0) mB[2] = __ldg(P.obs_meas + o * MP + 2);
        if (P.obs_w) wB = __ldg(P.obs_w + o);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            tB[k] = __ldg(P.lmk_t + 3 * (size_t)lB + k);
            xB[k] = B.xl[3 * (size_t)lB + k]; // (k_backsub mirrors the reduced-system entries of kept landmarks into xl)
        }
    };
    if (ol < Oloc) {
        load_idx(ol);
        gather(ol);
        if (ol + stride < Oloc) load_idx(ol + stride);
    }
    if constexpr (SMEM) mbar_wait(&bar, 0);
    for (; ol < Oloc; ol += stride) {
        const int fc = fcB;
        const double p[3] = {tB[0] + xB[0], tB[1] + xB[1], tB[2] + xB[2]};
        const double meas[3] = {mB[0], mB[1], mB[2]};
        double w = wB;
        double r[2], Jp[12], Jl[6];
        const double *row = fct + (size_t)fc * fstride;
        if (!P.obs_w) w = row[30];
        // request the operands of the next tile and the indices of the one after (every value of the previous requests has
        // been consumed above: ptxas tracks loads with a few counting scoreboards, a later wait on an old request would
        // also wait for the new ones)
        if (EARLY) {
            if (ol + stride < Oloc) gather(ol + stride);
            if (ol + 2 * stride < Oloc) load_idx(ol + 2 * stride);
        }
        eval_visual<KIND>(row, P.K + 4 * (fc % P.C), w, p, meas, r, Jp, Jl);
        if (!EARLY) {
            if (ol + stride < Oloc) gather(ol + stride);
            if (ol + 2 * stride < Oloc) load_idx(ol + 2 * stride);
        }
        B.r[ol] = r[0];
        B.r[(size_t)OC + ol] = r[1];
#pragma unroll
        for (int k = 0; k < 12; k++) B.Jp[(size_t)k * OC + ol] = Jp[k];
#pragma unroll
        for (int k = 0; k < 6; k++) B.Jl[(size_t)k * OC + ol] = Jl[k];
        csum += r[0] * r[0] + r[1] * r[1];
    }
    csum = warp_sum(csum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = csum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int i = 0; i < LIN_THREADS / 32; i++) s += red[i];
        atomicAdd(&acc->cost[b], 0.5 * s);
    }
}

// =====================================================================================================================
// non-visual factors: IMUFactor (a7), IMUBiasFactor (a8), PosePriordx (a9), MarginalizationFactor (a10)
// =====================================================================================================================
// Upper-triangular sqrt information U with U^T U = cov^-1 (residuals.hpp:151-154), once per uploaded window.
// One warp per IMU pair (4 pairs per CTA); lane j owns column j of the 9 x 18 tableau [cov | I], so every row operation of
// the elimination is one step for the warp.  Each element sees exactly the operations, in the order, of the sequential
// algorithm (the result is bit-identical to it); one thread per pair took 44 us per upload.
__global__ void __launch_bounds__(128) k_imu_inf_sqrt(const DevProblem *__restrict__ Pg) {
    const DevProblem &P = *Pg; // device-resident problem description: the launch parameters do not depend on the window (one CUDA graph serves them all)
    __shared__ double ms[4][9][19], Ls[4][9][9];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int p = blockIdx.x * 4 + wib;
    if (p >= P.P) return;
    double(*m)[19] = ms[wib];
    double(*Lm)[9] = Ls[wib];
    const double *cov = P.imu_cov + 81 * (size_t)p;
    const int j = lane; // my column (lanes 18.. idle)
    if (j < 18)
        for (int i = 0; i < 9; i++) m[i][j] = j < 9 ? cov[i * 9 + j] : (i == j - 9 ? 1.0 : 0.0);
    __syncwarp();
    // Gauss-Jordan with partial pivoting (what Eigen's PartialPivLU-based inverse() amounts to)
    for (int c = 0; c < 9; c++) {
        int piv = c;
        double best = fabs(m[c][c]);
        for (int r = c + 1; r < 9; r++)
            if (fabs(m[r][c]) > best) {
                best = fabs(m[r][c]);
                piv = r;
            }
        __syncwarp();
        if (piv != c && j < 18) {
            double t = m[c][j];
            m[c][j] = m[piv][j];
            m[piv][j] = t;
        }
        __syncwarp();
        const double inv = 1.0 / m[c][c];
        double f[9];
        for (int r = c + 1; r < 9; r++) f[r] = m[r][c] * inv;
        __syncwarp();
        if (j >= c && j < 18)
            for (int r = c + 1; r < 9; r++) m[r][j] -= f[r] * m[c][j];
        __syncwarp();
    }
    for (int c = 8; c >= 0; c--) {
        const double inv = 1.0 / m[c][c];
        __syncwarp();
        if (j < 18) m[c][j] *= inv;
        __syncwarp();
        double f[9];
        for (int r = 0; r < c; r++) f[r] = m[r][c];
        __syncwarp();
        if (j < 18)
            for (int r = 0; r < c; r++) m[r][j] -= f[r] * m[c][j];
        __syncwarp();
    }
    // Cholesky of the inverse: L L^T, store U = L^T; lane i owns row i
    const int i = lane;
    if (i < 9)
        for (int q = 0; q < 9; q++) Lm[i][q] = 0.0;
    __syncwarp();
    for (int q = 0; q < 9; q++) {
        double s = m[q][9 + q];
        for (int k = 0; k < q; k++) s -= Lm[q][k] * Lm[q][k];
        const double d = sqrt(s);
        __syncwarp();
        if (i == q) Lm[q][q] = d;
        if (i > q && i < 9) {
            double t = m[i][9 + q];
            for (int k = 0; k < q; k++) t -= Lm[i][k] * Lm[q][k];
            Lm[i][q] = t / d;
        }
        __syncwarp();
    }
    double *U = P.imu_inf_sqrt + 81 * (size_t)p;
    for (int e = lane; e < 81; e += 32) U[e] = Lm[e % 9][e / 9];
}

// current value of a frame's parameter blocks
SDV_DEV void frame_params(const DevProblem &P, const double *xp, int f, double *dpose, double *dv, double *dba, double *dbg) {
    int pc = P.pose_col[f], vc = P.vb_col[f];
#pragma unroll
    for (int k = 0; k < 6; k++) dpose[k] = pc >= 0 ? xp[pc + k] : 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        dv[k] = vc >= 0 ? xp[vc + k] : 0.0;
        dba[k] = vc >= 0 ? xp[vc + 3 + k] : 0.0;
        dbg[k] = vc >= 0 ? xp[vc + 6 + k] : 0.0;
    }
}

SDV_DEV void load_RT(const double *T, double *R, double *t) {
    R[0] = T[0]; R[1] = T[1]; R[2] = T[2];
    R[3] = T[4]; R[4] = T[5]; R[5] = T[6];
    R[6] = T[8]; R[7] = T[9]; R[8] = T[10];
    t[0] = T[3]; t[1] = T[7]; t[2] = T[11];
}

// whether a residual block belongs to the reduced program (at least one non-constant parameter block)
SDV_DEV bool imu_active(const DevProblem &P, int i, int j) {
    return P.pose_col[i] >= 0 || P.pose_col[j] >= 0 || P.vb_col[i] >= 0 || P.vb_col[j] >= 0;
}
SDV_DEV bool bias_active(const DevProblem &P, int i, int j) { return P.vb_col[i] >= 0 || P.vb_col[j] >= 0; }

constexpr int FAC_WARPS = 4;

// One warp per IMU pair; then one thread per pose prior; then the dense prior residual.  All lanes redundantly form
// the 3x3 building blocks, assemble the unwhitened 9x25 [J | r] in shared memory, and lane c whitens column c.
__global__ void __launch_bounds__(FAC_WARPS * 32) k_lin_factors(const DevProblem *__restrict__ Pg, LinBuf B0, LinBuf B1, const LMState *st, Accum *acc,
                                                                int which) {
    const DevProblem &P = *Pg; // device-resident problem description: the launch parameters do not depend on the window (one CUDA graph serves them all)
    if (st->status != 0) return;
    if (which == -2 && !st->step_valid) return;
    if (P.rank != 0) return; // non-visual factors live on rank 0 only
    int b = which >= 0 ? which : (which == -1 ? st->cur : 1 - st->cur);
    const LinBuf &B = b ? B1 : B0;
    __shared__ double Ju[FAC_WARPS][9 * 25];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * FAC_WARPS + wib;
    const int nw = gridDim.x * FAC_WARPS;
    double cost = 0.0, fcost = 0.0;
    // ---- IMU factors
    for (int p = gw; p < P.P; p += nw) {
        int i = P.imu_i[p], j = P.imu_j[p];
        double dxi[6], dxj[6], dvi[3], dvj[3], dba[3], dbg[3], t3a[3], t3b[3];
        frame_params(P, B.xp, i, dxi, dvi, dba, dbg);
        frame_params(P, B.xp, j, dxj, dvj, t3a, t3b);
        double Rib[9], tib[3], Rjb[9], tjb[3];
        load_RT(P.T_f_w + 12 * i, Rib, tib);
        load_RT(P.T_f_w + 12 * j, Rjb, tjb);
        double dRi[9], dRj[9], Ri[9], Rj[9], ti[3], tj[3], tmp[3];
        exp_so3(dxi, dRi);
        exp_so3(dxj, dRj);
        mat3_mul(Rib, dRi, Ri);
        mat3_mul(Rjb, dRj, Rj);
        mat3_vec(Rib, dxi + 3, tmp);
        ti[0] = tmp[0] + tib[0]; ti[1] = tmp[1] + tib[1]; ti[2] = tmp[2] + tib[2];
        mat3_vec(Rjb, dxj + 3, tmp);
        tj[0] = tmp[0] + tjb[0]; tj[1] = tmp[1] + tjb[1]; tj[2] = tmp[2] + tjb[2];
        double vi[3], vj[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            vi[k] = P.v[3 * i + k] + dvi[k];
            vj[k] = P.v[3 * j + k] + dvj[k];
        }
        const double dt = P.imu_dt[p];
        const double g[3] = {0.0, 0.0, -9.81};
        const double *DR = P.imu_dR + 9 * (size_t)p, *JRg = P.imu_J_dR_bg + 9 * (size_t)p;
        const double *Jva = P.imu_J_dv_ba + 9 * (size_t)p, *Jvg = P.imu_J_dv_bg + 9 * (size_t)p;
        const double *Jpa = P.imu_J_dp_ba + 9 * (size_t)p, *Jpg = P.imu_J_dp_bg + 9 * (size_t)p;
        // dR = (DeltaR * Exp(J_dR_bg dbg))^T R_i R_j^T      (residuals.hpp:157-158)
        double phi[3], Eb[9], DRc[9], RiRjT[9], dR[9], r_dr[3];
        mat3_vec(JRg, dbg, phi);
        exp_so3(phi, Eb);
        mat3_mul(DR, Eb, DRc);
        mat3_mulT(Ri, Rj, RiRjT);
        matT3_mul(DRc, RiRjT, dR);
        log_so3(dR, r_dr);
        // r_dv, r_dp (residuals.hpp:160-164)
        double a[3], pj[3], pi[3], bvec[3], c[3], r_dv[3], r_dp[3];
#pragma unroll
        for (int k = 0; k < 3; k++) a[k] = vj[k] - vi[k] - g[k] * dt;
        matT3_vec(Rj, tj, pj);
        matT3_vec(Ri, ti, pi);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            pj[k] = -pj[k];
            pi[k] = -pi[k];
            bvec[k] = pj[k] - vi[k] * dt - 0.5 * g[k] * dt * dt; // argument of the skew in J_dTfi block(6,0)
            c[k] = pj[k] - pi[k] - vi[k] * dt - 0.5 * g[k] * dt * dt;
        }
        double Ra[3], Rc[3], cv[3], cp[3];
        mat3_vec(Ri, a, Ra);
        mat3_vec(Ri, c, Rc);
        {
            double t1[3], t2[3];
            mat3_vec(Jvg, dbg, t1);
            mat3_vec(Jva, dba, t2);
#pragma unroll
            for (int k = 0; k < 3; k++) cv[k] = P.imu_dv[3 * (size_t)p + k] + t1[k] + t2[k];
            mat3_vec(Jpg, dbg, t1);
            mat3_vec(Jpa, dba, t2);
#pragma unroll
            for (int k = 0; k < 3; k++) cp[k] = P.imu_dp[3 * (size_t)p + k] + t1[k] + t2[k];
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            r_dv[k] = Ra[k] - cv[k];
            r_dp[k] = Rc[k] - cp[k];
        }
        // ---- 3x3 blocks of the unwhitened Jacobian (residuals.hpp:174-237)
        double Jri[9], Jrj[9], JrInv[9], tmpM[9], tmpN[9], S[9];
        right_jacobian(dxi, Jri);
        right_jacobian(dxj, Jrj);
        right_jacobian(r_dr, tmpM);
        inverse3(tmpM, JrInv);
        double B00[9], B30[9], B60[9], C00[9], C60[9], C63[9], D00[9];
        mat3_mul(JrInv, Rj, tmpM);          // Jr^-1(r_dr) R_j
        mat3_mul(tmpM, Jri, B00);           // pose_i rot -> r_dr
        mat3_mul(tmpM, Jrj, C00);           // (negated below) pose_j rot -> r_dr
        skew3(a, S);
        mat3_mul(Ri, S, tmpN);
        mat3_mul(tmpN, Jri, B30);           // (negated) pose_i rot -> r_dv
        skew3(bvec, S);
        mat3_mul(Ri, S, tmpN);
        mat3_mul(tmpN, Jri, B60);           // (negated) pose_i rot -> r_dp
        skew3(tj, S);
        mat3_mul(RiRjT, S, tmpN);
        mat3_mul(tmpN, Rj, tmpM);
        mat3_mul(tmpM, Jrj, C60);           // (negated) pose_j rot -> r_dp
        mat3_mulT(Ri, dRj, C63);            // (negated) pose_j trans -> r_dp :  R_i exp(w_j)^T
        {
            double Jrb[9];
            right_jacobian(phi, Jrb);
            mat3_mulT(JrInv, dR, tmpM);     // Jr^-1 dR^T
            mat3_mul(tmpM, Jrb, tmpN);
            mat3_mul(tmpN, JRg, D00);       // (negated) bg -> r_dr
        }
        // ---- assemble [J | r] (9 x 25) in shared memory
        double *M = Ju[wib];
        for (int k = lane; k < 225; k += 32) M[k] = 0.0;
        __syncwarp();
        if (lane < 9) {
            int rr = lane / 3, cc = lane % 3, k = lane;
            M[(0 + rr) * 25 + 0 + cc] = B00[k];
            M[(3 + rr) * 25 + 0 + cc] = -B30[k];
            M[(6 + rr) * 25 + 0 + cc] = -B60[k];
            M[(6 + rr) * 25 + 3 + cc] = Rib[k];
            M[(0 + rr) * 25 + 6 + cc] = -C00[k];
            M[(6 + rr) * 25 + 6 + cc] = -C60[k];
            M[(6 + rr) * 25 + 9 + cc] = -C63[k];
            M[(3 + rr) * 25 + 12 + cc] = -Ri[k];
            M[(6 + rr) * 25 + 12 + cc] = -Ri[k] * dt;
            M[(3 + rr) * 25 + 15 + cc] = Ri[k];
            M[(3 + rr) * 25 + 18 + cc] = -Jva[k];
            M[(6 + rr) * 25 + 18 + cc] = -Jpa[k];
            M[(0 + rr) * 25 + 21 + cc] = -D00[k];
            M[(3 + rr) * 25 + 21 + cc] = -Jvg[k];
            M[(6 + rr) * 25 + 21 + cc] = -Jpg[k];
        }
        if (lane < 3) {
            M[(0 + lane) * 25 + 24] = r_dr[lane];
            M[(3 + lane) * 25 + 24] = r_dv[lane];
            M[(6 + lane) * 25 + 24] = r_dp[lane];
        }
        __syncwarp();
        // ---- whiten: column c of U * [J | r]
        if (lane < 25) {
            const double *U = P.imu_inf_sqrt + 81 * (size_t)p;
            double col[9], out[9];
#pragma unroll
            for (int k = 0; k < 9; k++) col[k] = M[k * 25 + lane];
#pragma unroll
            for (int r = 0; r < 9; r++) {
                double s = 0;
#pragma unroll
                for (int k = 0; k < 9; k++)
                    if (k >= r) s += U[r * 9 + k] * col[k];
                out[r] = s;
            }
            if (lane < 24) {
#pragma unroll
                for (int r = 0; r < 9; r++) B.imu_J[(size_t)p * 216 + r * 24 + lane] = out[r];
            } else {
                double s = 0;
#pragma unroll
                for (int r = 0; r < 9; r++) {
                    B.imu_r[(size_t)p * 9 + r] = out[r];
                    s += out[r] * out[r];
                }
                if (imu_active(P, i, j)) cost += s;
                else fcost += s;
            }
        }
        __syncwarp();
        // ---- bias random-walk factor (residuals.hpp:252-266)
        if (lane < 6) {
            bool is_ba = lane < 3;
            int k = lane % 3;
            double sig = is_ba ? P.imu_sigma_ba[p] : P.imu_sigma_bg[p];
            double w = 1.0 / sqrt(dt * sig * sig);
            double bi = is_ba ? P.ba[3 * i + k] + dba[k] : P.bg[3 * i + k] + dbg[k];
            double bj = is_ba ? P.ba[3 * j + k] + t3a[k] : P.bg[3 * j + k] + t3b[k];
            double r = w * (bj - bi);
            B.bias_r[(size_t)p * 6 + lane] = r;
            if (bias_active(P, i, j)) cost += r * r;
            else fcost += r * r;
        }
    }
    // ---- pose priors (residuals.hpp:607-628), one thread each
    if (P.has_prior) {
        int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
        for (int f = tid; f < P.F; f += nt) {
            if (!P.has_prior[f]) continue;
            double dx[6], d3[3], e3[3], f3[3];
            frame_params(P, B.xp, f, dx, d3, e3, f3);
            double Rb[9], tb[3], Rp[9], tp[3], dR[9], R[9], t[3], tmp[3];
            load_RT(P.T_f_w + 12 * f, Rb, tb);
            load_RT(P.T_prior + 12 * f, Rp, tp);
            exp_so3(dx, dR);
            mat3_mul(Rb, dR, R);
            mat3_vec(Rb, dx + 3, tmp);
            t[0] = tmp[0] + tb[0]; t[1] = tmp[1] + tb[1]; t[2] = tmp[2] + tb[2];
            // T * T_prior^-1 : rotation R Rp^T, translation t - R Rp^T tp
            double RRt[9], w[3], c3[3], Rc[3], e[6];
            mat3_mulT(R, Rp, RRt);
            log_so3(RRt, w);
            matT3_vec(Rp, tp, c3); // Rp^T tp
            mat3_vec(R, c3, Rc);
            e[0] = w[0]; e[1] = w[1]; e[2] = w[2];
            e[3] = t[0] - Rc[0]; e[4] = t[1] - Rc[1]; e[5] = t[2] - Rc[2];
            const double *sw = P.inf_prior + 6 * f;
            double Jrw[9], JrwInv[9], Jrd[9], M1[9], J00[9], J30[9], S[9];
            right_jacobian(w, Jrw);
            inverse3(Jrw, JrwInv);
            right_jacobian(dx, Jrd);
            mat3_mul(JrwInv, Rp, M1);
            mat3_mul(M1, Jrd, J00);
            skew3(c3, S);
            mat3_mul(R, S, M1);
            mat3_mul(M1, Jrd, J30);
            double *J = B.prior_J + 36 * (size_t)f;
            double *r = B.prior_r + 6 * (size_t)f;
            double s = 0;
            for (int rr = 0; rr < 3; rr++)
                for (int cc = 0; cc < 3; cc++) {
                    J[rr * 6 + cc] = sw[rr] * J00[rr * 3 + cc];
                    J[rr * 6 + 3 + cc] = 0.0;
                    J[(3 + rr) * 6 + cc] = sw[3 + rr] * J30[rr * 3 + cc];
                    J[(3 + rr) * 6 + 3 + cc] = sw[3 + rr] * Rb[rr * 3 + cc];
                }
            for (int k = 0; k < 6; k++) {
                r[k] = sw[k] * e[k];
                s += r[k] * r[k];
            }
            if (P.pose_col[f] >= 0) cost += s; // all-constant blocks go to Ceres' fixed_cost, not the cost
            else fcost += s;
        }
    }
    // ---- dense marginalisation prior: r = r0 + J_m dx  (marginalization.hpp:113-148)
    if (P.mp_nfull > 0) {
        int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
        bool any_free = false;
        for (int k = 0; k < P.mp_nmap; k++) any_free |= P.mp_dst_col[k] >= 0;
        for (int row = tid; row < P.mp_nfull; row += nt) {
            double s = P.mp_r0[row];
            const double *Jr = P.mp_J + (size_t)row * P.mp_n;
            for (int k = 0; k < P.mp_nmap; k++) {
                int dc = P.mp_dst_col[k];
                if (dc >= 0) s += Jr[P.mp_src_col[k]] * B.xp[dc];
            }
            B.mp_r[row] = s;
            if (any_free) cost += s * s;
            else fcost += s * s;
        }
    }
    // ---- sparsified prior: IMUPriordx (residuals.hpp:634-700), Landmark3DPrior (:506-526), LandmarkToLandmark (:528-559)
    if (blockIdx.x == 0 && threadIdx.x == 0 && P.sp_has_imu) {
        const int f = P.sp_frame;
        const double *blob = P.sp_blob;
        double dx[6], dv[3], dba[3], dbg[3];
        frame_params(P, B.xp, f, dx, dv, dba, dbg);
        double Rb[9], tb[3], Rp[9], tp[3], dR[9], R[9], t[3], tmp[3];
        load_RT(P.T_f_w + 12 * f, Rb, tb);
        load_RT(blob, Rp, tp);
        exp_so3(dx, dR);
        mat3_mul(Rb, dR, R);
        mat3_vec(Rb, dx + 3, tmp);
        t[0] = tmp[0] + tb[0]; t[1] = tmp[1] + tb[1]; t[2] = tmp[2] + tb[2];
        double RRt[9], w[3], c3[3], Rc[3], e[15];
        mat3_mulT(R, Rp, RRt);
        log_so3(RRt, w);
        matT3_vec(Rp, tp, c3);
        mat3_vec(R, c3, Rc);
        for (int k = 0; k < 3; k++) {
            e[k] = w[k];
            e[3 + k] = t[k] - Rc[k];
            e[6 + k] = P.v[3 * f + k] + dv[k] - blob[12 + k];
            e[9 + k] = P.ba[3 * f + k] + dba[k] - blob[15 + k];
            e[12 + k] = P.bg[3 * f + k] + dbg[k] - blob[18 + k];
        }
        const double *SQ = blob + 21; // 15x15
        double Jrw[9], JrwInv[9], Jrd[9], M1[9], J00[9], J30[9], S3[9];
        right_jacobian(w, Jrw);
        inverse3(Jrw, JrwInv);
        right_jacobian(dx, Jrd);
        mat3_mul(JrwInv, Rp, M1);
        mat3_mul(M1, Jrd, J00);
        skew3(c3, S3);
        mat3_mul(R, S3, M1);
        mat3_mul(M1, Jrd, J30);
        double s = 0;
        for (int i = 0; i < 15; i++) {
            double ri = 0;
            for (int k = 0; k < 15; k++) ri += SQ[i * 15 + k] * e[k];
            B.sp_r[i] = ri;
            s += ri * ri;
            // pose block: sqrt_inf * [J00 0; J30 Rb; 0 0]
            for (int c = 0; c < 3; c++) {
                double a = 0, bb = 0;
                for (int k = 0; k < 3; k++) {
                    a += SQ[i * 15 + k] * J00[k * 3 + c] + SQ[i * 15 + 3 + k] * J30[k * 3 + c];
                    bb += SQ[i * 15 + 3 + k] * Rb[k * 3 + c];
                }
                B.sp_J[i * 15 + c] = a;
                B.sp_J[i * 15 + 3 + c] = bb;
            }
            // reference quirk (residuals.hpp:676-692): the v / ba / bg Jacobians are identity blocks, NOT whitened
            for (int c = 6; c < 15; c++) B.sp_J[i * 15 + c] = (i == c) ? 1.0 : 0.0;
        }
        if (P.pose_col[f] >= 0 || P.vb_col[f] >= 0) cost += s;
        else fcost += s;
    }
    if (blockIdx.x == 0 && threadIdx.x == 1 && P.sp_has_lmk) {
        const double *blob = P.sp_blob;
        double p[3];
        landmark_position(P, B, P.sp_lmk0, p);
        double d[3] = {p[0] - blob[246], p[1] - blob[247], p[2] - blob[248]};
        const double *SQ = blob + 249;
        for (int i = 0; i < 3; i++) {
            double ri = SQ[i * 3] * d[0] + SQ[i * 3 + 1] * d[1] + SQ[i * 3 + 2] * d[2];
            B.sp_r[15 + i] = ri;
            cost += ri * ri;
        }
    }
    {
        int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
        for (int k = tid; k < P.sp_nl2l; k += nt) {
            double pa[3], pb[3];
            landmark_position(P, B, P.sp_l2l_a[k], pa);
            landmark_position(P, B, P.sp_l2l_b[k], pb);
            double d[3] = {pa[0] - pb[0] - P.sp_l2l_delta[3 * k], pa[1] - pb[1] - P.sp_l2l_delta[3 * k + 1], pa[2] - pb[2] - P.sp_l2l_delta[3 * k + 2]};
            const double *SQ = P.sp_l2l_sqrt + 9 * k;
            for (int i = 0; i < 3; i++) {
                double ri = SQ[i * 3] * d[0] + SQ[i * 3 + 1] * d[1] + SQ[i * 3 + 2] * d[2];
                B.sp_r[18 + 3 * k + i] = ri;
                cost += ri * ri;
            }
        }
    }
    cost = warp_sum(cost);
    if (lane == 0 && cost != 0.0) atomicAdd(&acc->cost[b], 0.5 * cost);
    fcost = warp_sum(fcost);
    if (lane == 0 && fcost != 0.0 && which == 0) atomicAdd(&acc->fixed_cost, 0.5 * fcost);
}

// PoseToLandmarkFactor (residuals.hpp:561-599): 3 residuals on (kept frame pose, landmark). The landmark stays in the
// eliminated set, so each factor is written as TWO pseudo-observations (rows 0-1 and row 2 + a zero row) into the r / J
// planes of the owning rank and flows through the same per-landmark Schur machinery as the visual factors.
__global__ void k_lin_p2l(const DevProblem *__restrict__ Pg, LinBuf B0, LinBuf B1, const LMState *st, Accum *acc, int which) {
    const DevProblem &P = *Pg; // device-resident problem description: the launch parameters do not depend on the window (one CUDA graph serves them all)
    if (st->status != 0) return;
    if (which == -2 && !st->step_valid) return;
    int b = which >= 0 ? which : (which == -1 ? st->cur : 1 - st->cur);
    const LinBuf &B = b ? B1 : B0;
    const int OC = P.Ocap;
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    double cost = 0.0;
    if (k < P.sp_np2l && P.sp_p2l_plane[k] >= 0) {
        const int f = P.sp_frame, l = P.sp_p2l_lmk[k], pl = P.sp_p2l_plane[k];
        double dx[6], d3[3], e3[3], f3[3];
        frame_params(P, B.xp, f, dx, d3, e3, f3);
        double Rb[9], tb[3], dR[9], R[9], t[3], tmp[3], p[3];
        load_RT(P.T_f_w + 12 * f, Rb, tb);
        exp_so3(dx, dR);
        mat3_mul(Rb, dR, R);
        mat3_vec(Rb, dx + 3, tmp);
        t[0] = tmp[0] + tb[0]; t[1] = tmp[1] + tb[1]; t[2] = tmp[2] + tb[2];
        landmark_position(P, B, l, p);
        double Tp[3];
        mat3_vec(R, p, Tp);
        const double *SQ = P.sp_p2l_sqrt + 9 * k, *dl = P.sp_p2l_delta + 3 * k;
        double e[3] = {Tp[0] + t[0] - dl[0], Tp[1] + t[1] - dl[1], Tp[2] + t[2] - dl[2]};
        double r[3], SR[9], Jrot[9], Jtr[9], Jl[9], S3[9], Jr[9], M1[9], M2[9];
        mat3_vec(SQ, e, r);
        mat3_mul(SQ, Rb, SR);          // sqrt_inf * R_base
        skew3(p, S3);
        right_jacobian(dx, Jr);
        mat3_mul(dR, S3, M1);
        mat3_mul(M1, Jr, M2);          // dR [t]x Jr(dw)
        mat3_mul(SR, M2, Jrot);
#pragma unroll
        for (int q = 0; q < 9; q++) {
            Jrot[q] = -Jrot[q];
            Jtr[q] = SR[q];
        }
        mat3_mul(SQ, R, Jl);           // sqrt_inf * R
        // pseudo-observation a: rows 0,1 ; pseudo-observation b: row 2 and a zero row
        B.r[pl] = r[0];
        B.r[(size_t)OC + pl] = r[1];
        B.r[pl + 1] = r[2];
        B.r[(size_t)OC + pl + 1] = 0.0;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            B.Jp[(size_t)c * OC + pl] = Jrot[c];
            B.Jp[(size_t)(3 + c) * OC + pl] = Jtr[c];
            B.Jp[(size_t)(6 + c) * OC + pl] = Jrot[3 + c];
            B.Jp[(size_t)(9 + c) * OC + pl] = Jtr[3 + c];
            B.Jl[(size_t)c * OC + pl] = Jl[c];
            B.Jl[(size_t)(3 + c) * OC + pl] = Jl[3 + c];
            B.Jp[(size_t)c * OC + pl + 1] = Jrot[6 + c];
            B.Jp[(size_t)(3 + c) * OC + pl + 1] = Jtr[6 + c];
            B.Jp[(size_t)(6 + c) * OC + pl + 1] = 0.0;
            B.Jp[(size_t)(9 + c) * OC + pl + 1] = 0.0;
            B.Jl[(size_t)c * OC + pl + 1] = Jl[6 + c];
            B.Jl[(size_t)(3 + c) * OC + pl + 1] = 0.0;
        }
        cost = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    }
    cost = warp_sum(cost);
    if ((threadIdx.x & 31) == 0 && cost != 0.0) atomicAdd(&acc->cost[b], 0.5 * cost);
}

// =====================================================================================================================
// per-landmark accumulation shared by k_schur and k_backsub
// =====================================================================================================================
SDV_DEV int tri_idx(int i, int j) { return i * (i + 1) / 2 + j; } // i >= j

struct SlotAcc {
    double W[18]; // Jp^T Jl  (6x3)
    double H[21]; // lower triangle of Jp^T Jp
    double gp[6]; // Jp^T r
};

// LM damping of one column in UNSCALED variables: clamp(s^2 c, lo, hi) / (radius s^2)
SDV_DEV double lm_damping(double c, double s, double radius, const SolverOpts &o) {
    double d = fmin(fmax(s * s * c, o.min_diag), o.max_diag);
    return d / (radius * s * s);
}

constexpr int MAX_SLOTS = 32; // distinct keyframes one landmark may be seen from (checked at upload)

// Layout of the reduced-system buffer Sb: rows [0, n_pad) = S (lower triangle used), row n_pad = g (right-hand side),
// row n_pad+1 = diag(J^T J) of the reduced columns (before damping), row n_pad+2 = raw gradient of the reduced columns.
// J^T J / J^T r of the non-visual factors into the reduced system (rank 0 only). One warp per IMU pair.
__global__ void __launch_bounds__(FAC_WARPS * 32) k_assemble_factors(const DevProblem *__restrict__ Pg, LinBuf B0, LinBuf B1, const LMState *st, double *Sb) {
    const DevProblem &P = *Pg; // device-resident problem description: the launch parameters do not depend on the window (one CUDA graph serves them all)
    if (st->status != 0 || P.rank != 0) return;
    const LinBuf &B = st->cur ? B1 : B0;
    __shared__ double Js[FAC_WARPS][9 * 25];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * FAC_WARPS + wib, nw = gridDim.x * FAC_WARPS;
    const int ld = P.ld;
    double *g = Sb + (size_t)P.n_pad * ld, *cdiag = g + ld, *graw = cdiag + ld;
    for (int p = gw; p < P.P; p += nw) {
        int i = P.imu_i[p], j = P.imu_j[p];
        // column of each of the 8 three-column groups: pose_i rot, pose_i trans, pose_j rot, pose_j trans, v_i, v_j, ba_i, bg_i
        int pci = P.pose_col[i], pcj = P.pose_col[j], vci = P.vb_col[i], vcj = P.vb_col[j];
        int gcol[8] = {pci, pci >= 0 ? pci + 3 : -1, pcj, pcj >= 0 ? pcj + 3 : -1, vci, vcj, vci >= 0 ? vci + 3 : -1, vci >= 0 ? vci + 6 : -1};
        double *M = Js[wib];
        for (int k = lane; k < 216; k += 32) M[(k / 24) * 25 + (k % 24)] = B.imu_J[(size_t)p * 216 + k];
        if (lane < 9) M[lane * 25 + 24] = B.imu_r[(size_t)p * 9 + lane];
        __syncwarp();
        for (int e = lane; e < 24 * 25; e += 32) {
            int a = e / 25, bcol = e - a * 25;
            int ra = gcol[a / 3];
            if (ra < 0) continue;
            ra += a % 3;
            double s = 0;
#pragma unroll
            for (int r = 0; r < 9; r++) s += M[r * 25 + a] * M[r * 25 + bcol];
            if (bcol == 24) {
                atomicAdd(&g[ra], s);
                atomicAdd(&graw[ra], s);
            } else {
                int rb = gcol[bcol / 3];
                if (rb < 0) continue;
                rb += bcol % 3;
                if (ra >= rb) atomicAdd(&Sb[(size_t)ra * ld + rb], s);
                if (a == bcol) atomicAdd(&cdiag[ra], s);
            }
        }
        __syncwarp();
        // bias random walk: J = -w on block i, +w on block j
        if (lane < 6) {
            bool is_ba = lane < 3;
            int k = lane % 3;
            double sig = is_ba ? P.imu_sigma_ba[p] : P.imu_sigma_bg[p];
            double w = 1.0 / sqrt(P.imu_dt[p] * sig * sig);
            double r = B.bias_r[(size_t)p * 6 + lane];
            int ci = vci >= 0 ? vci + (is_ba ? 3 : 6) + k : -1;
            int cj = vcj >= 0 ? vcj + (is_ba ? 3 : 6) + k : -1;
            if (ci >= 0) {
                atomicAdd(&Sb[(size_t)ci * ld + ci], w * w);
                atomicAdd(&cdiag[ci], w * w);
                atomicAdd(&g[ci], -w * r);
                atomicAdd(&graw[ci], -w * r);
            }
            if (cj >= 0) {
                atomicAdd(&Sb[(size_t)cj * ld + cj], w * w);
                atomicAdd(&cdiag[cj], w * w);
                atomicAdd(&g[cj], w * r);
                atomicAdd(&graw[cj], w * r);
            }
            if (ci >= 0 && cj >= 0) {
                if (ci > cj) atomicAdd(&Sb[(size_t)ci * ld + cj], -w * w);
                else atomicAdd(&Sb[(size_t)cj * ld + ci], -w * w);
            }
        }
    }
    // pose priors
    if (P.has_prior) {
        int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
        for (int f = tid; f < P.F; f += nt) {
            int pc = P.pose_col[f];
            if (!P.has_prior[f] || pc < 0) continue;
            const double *J = B.prior_J + 36 * (size_t)f, *r = B.prior_r + 6 * (size_t)f;
            for (int a = 0; a < 6; a++) {
                double gs = 0;
                for (int q = 0; q < 6; q++) gs += J[q * 6 + a] * r[q];
                atomicAdd(&g[pc + a], gs);
                atomicAdd(&graw[pc + a], gs);
                for (int bb = 0; bb <= a; bb++) {
                    double s = 0;
                    for (int q = 0; q < 6; q++) s += J[q * 6 + a] * J[q * 6 + bb];
                    atomicAdd(&Sb[(size_t)(pc + a) * ld + pc + bb], s);
                    if (a == bb) atomicAdd(&cdiag[pc + a], s);
                }
            }
        }
    }
    // dense marginalisation prior: H_m = J_m^T J_m (constant, precomputed), g_m = J_m^T r0 + H_m dx
    if (P.mp_nfull > 0) {
        int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
        const int nm = P.mp_nmap;
        for (int e = tid; e < nm * nm; e += nt) {
            int a = e / nm, bb = e - a * nm;
            int ra = P.mp_dst_col[a], rb = P.mp_dst_col[bb];
            if (ra < 0 || rb < 0 || ra < rb) continue;
            double h = P.mp_H[(size_t)a * nm + bb];
            atomicAdd(&Sb[(size_t)ra * ld + rb], h);
            if (a == bb) atomicAdd(&cdiag[ra], h);
        }
        for (int a = tid; a < nm; a += nt) {
            int ra = P.mp_dst_col[a];
            if (ra < 0) continue;
            double s = P.mp_g0[a];
            for (int bb = 0; bb < nm; bb++) {
                int rb = P.mp_dst_col[bb];
                if (rb >= 0) s += P.mp_H[(size_t)a * nm + bb] * B.xp[rb];
            }
            atomicAdd(&g[ra], s);
            atomicAdd(&graw[ra], s);
        }
    }
}

// (k_assemble_factors continues in k_assemble_sparse for the sparsified prior)
__global__ void k_assemble_sparse(const DevProblem *__restrict__ Pg, LinBuf B0, LinBuf B1, const LMState *st, double *Sb) {
    const DevProblem &P = *Pg; // device-resident problem description: the launch parameters do not depend on the window (one CUDA graph serves them all)
    if (st->status != 0 || P.rank != 0) return;
    const LinBuf &B = st->cur ? B1 : B0;
    const int ld = P.ld;
    double *g = Sb + (size_t)P.n_pad * ld, *cdiag = g + ld, *graw = cdiag + ld;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    if (P.sp_has_imu) {
        const int f = P.sp_frame, pc = P.pose_col[f], vc = P.vb_col[f];
        for (int e = tid; e < 15 * 16; e += nt) {
            int a = e / 16, bcol = e - a * 16;
            int ra = a < 6 ? (pc >= 0 ? pc + a : -1) : (vc >= 0 ? vc + a - 6 : -1);
            if (ra < 0) continue;
            double s = 0;
            if (bcol == 15) {
                for (int i = 0; i < 15; i++) s += B.sp_J[i * 15 + a] * B.sp_r[i];
                atomicAdd(&g[ra], s);
                atomicAdd(&graw[ra], s);
            } else {
                int rb = bcol < 6 ? (pc >= 0 ? pc + bcol : -1) : (vc >= 0 ? vc + bcol - 6 : -1);
                if (rb < 0 || ra < rb) continue;
                for (int i = 0; i < 15; i++) s += B.sp_J[i * 15 + a] * B.sp_J[i * 15 + bcol];
                atomicAdd(&Sb[(size_t)ra * ld + rb], s);
                if (a == bcol) atomicAdd(&cdiag[ra], s);
            }
        }
    }
    if (P.sp_has_lmk && tid == 0) {
        const int dc = P.lmk_col[P.sp_lmk0];
        const double *SQ = P.sp_blob + 249, *r = B.sp_r + 15;
        for (int a = 0; a < 3; a++) {
            double gs = 0;
            for (int i = 0; i < 3; i++) gs += SQ[i * 3 + a] * r[i];
            atomicAdd(&g[dc + a], gs);
            atomicAdd(&graw[dc + a], gs);
            for (int bb = 0; bb <= a; bb++) {
                double s = 0;
                for (int i = 0; i < 3; i++) s += SQ[i * 3 + a] * SQ[i * 3 + bb];
                atomicAdd(&Sb[(size_t)(dc + a) * ld + dc + bb], s);
                if (a == bb) atomicAdd(&cdiag[dc + a], s);
            }
        }
    }
    for (int k = tid; k < P.sp_nl2l; k += nt) {
        const int ca = P.lmk_col[P.sp_l2l_a[k]], cb = P.lmk_col[P.sp_l2l_b[k]];
        const double *SQ = P.sp_l2l_sqrt + 9 * k, *r = B.sp_r + 18 + 3 * k;
        for (int a = 0; a < 3; a++) {
            double gs = 0;
            for (int i = 0; i < 3; i++) gs += SQ[i * 3 + a] * r[i];
            atomicAdd(&g[ca + a], gs);
            atomicAdd(&graw[ca + a], gs);
            atomicAdd(&g[cb + a], -gs);
            atomicAdd(&graw[cb + a], -gs);
            for (int bb = 0; bb < 3; bb++) {
                double s = 0;
                for (int i = 0; i < 3; i++) s += SQ[i * 3 + a] * SQ[i * 3 + bb];
                if (bb <= a) {
                    atomicAdd(&Sb[(size_t)(ca + a) * ld + ca + bb], s);
                    atomicAdd(&Sb[(size_t)(cb + a) * ld + cb + bb], s);
                }
                if (a == bb) {
                    atomicAdd(&cdiag[ca + a], s);
                    atomicAdd(&cdiag[cb + a], s);
                }
                // cross block J_a^T J_b = -M, stored in the lower triangle
                if (ca > cb) atomicAdd(&Sb[(size_t)(ca + a) * ld + cb + bb], -s);
                else atomicAdd(&Sb[(size_t)(cb + bb) * ld + ca + a], -s);
            }
        }
    }
}

// H_m = J_m^T J_m and g0 = J_m^T r0 restricted to the mapped columns; once per upload.
__global__ void k_prior_setup(const DevProblem *__restrict__ Pg) {
    const DevProblem &P = *Pg; // device-resident problem description: the launch parameters do not depend on the window (one CUDA graph serves them all)
    const int nm = P.mp_nmap;
    int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int e = tid; e < nm * nm + nm; e += nt) {
        if (e < nm * nm) {
            int a = e / nm, b = e - a * nm;
            int ca = P.mp_src_col[a], cb = P.mp_src_col[b];
            double s = 0;
            for (int r = 0; r < P.mp_nfull; r++) s += P.mp_J[(size_t)r * P.mp_n + ca] * P.mp_J[(size_t)r * P.mp_n + cb];
            P.mp_H[e] = s;
        } else {
            int a = e - nm * nm, ca = P.mp_src_col[a];
            double s = 0;
            for (int r = 0; r < P.mp_nfull; r++) s += P.mp_J[(size_t)r * P.mp_n + ca] * P.mp_r0[r];
            P.mp_g0[a] = s;
        }
    }
}

// =====================================================================================================================
// reduced system preparation: jacobi scaling (iteration 0), gradient test, LM damping, padding
// =====================================================================================================================
__global__ void __launch_bounds__(1024) k_sysprep(const DevProblem *__restrict__ Pg, LMState *st, Accum *acc, SolverOpts opt, double *Sb, double *scale_p,
                                                  double *damp_p, double *graw_p) {
    const DevProblem &P = *Pg; // device-resident problem description: the launch parameters do not depend on the window (one CUDA graph serves them all)
    if (st->status != 0) return;
    __shared__ double red[32];
    __shared__ int s_stop;
    const int ld = P.ld, n = P.n;
    double *g = Sb + (size_t)P.n_pad * ld, *cdiag = g + ld, *graw = cdiag + ld;
    const bool first = st->scaling_done == 0;
    const double radius = st->radius;
    if (threadIdx.x == 0) s_stop = 0;
    __syncthreads();
    // gradient tolerance (Ceres: iteration 0 and after every successful step)
    if (st->need_grad_check) {
        double m = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, fabs(graw[i]));
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            double mm = __longlong_as_double((long long)acc->grad_max_bits);
            for (int i = 0; i < (int)(blockDim.x >> 5); i++) mm = fmax(mm, red[i]);
            if (mm <= opt.gradient_tolerance) {
                st->status = 1 + 2; // SDV_TERM_GRADIENT_TOLERANCE
                st->iter -= 1;      // k_iter_begin already counted a step Ceres never starts
                s_stop = 1;
            }
            st->need_grad_check = 0;
        }
        __syncthreads();
        if (s_stop) return;
    }
    for (int i = threadIdx.x; i < P.n_pad; i += blockDim.x) {
        if (i < n) {
            double c = cdiag[i];
            double s = first ? (opt.jacobi_scaling ? 1.0 / (1.0 + sqrt(c)) : 1.0) : scale_p[i];
            if (first) scale_p[i] = s;
            double d = lm_damping(c, s, radius, opt);
            damp_p[i] = d;
            graw_p[i] = graw[i];
            Sb[(size_t)i * ld + i] += d;
        } else {
            Sb[(size_t)i * ld + i] = 1.0; // padding: identity block, zero right-hand side
            g[i] = 0.0;
            damp_p[i] = 0.0;
            graw_p[i] = 0.0;
        }
    }
    __syncthreads();
    // rows n_pad+1 .. n_pad+31 ride along through the factorisation as extra "rows below": keep them zero
    for (int i = threadIdx.x; i < 31 * ld; i += blockDim.x) cdiag[i] = 0.0;
    if (threadIdx.x == 0) {
        st->scaling_done = 1;
        acc->grad_max_bits = 0ull;
    }
}

// =====================================================================================================================
// dense Cholesky of the reduced system, right-looking, one launch per 32-column panel.
//   A  : (n_pad + 32) x ld, lower triangle; rows n_pad.. hold the right-hand side (row n_pad) so that the forward
//        substitution L y = g falls out of the panel solves.
//   Lo : same shape, receives L (and y^T in row n_pad).
// CTA roles for panel k (tiles are 32x32, T = n_pad/32):
//   blockIdx.x <  T-k+1           : panel CTA for tile row i = k + blockIdx.x (i == T is the right-hand-side row)
//   otherwise                      : update CTA for a trailing tile (i, j), k < j <= i <= T, j < T
// Every CTA re-factors A_kk (redundant, but it removes the inter-CTA dependency inside the launch).
// =====================================================================================================================
constexpr int CH_T = 32;
constexpr int CH_THREADS = 256;

SDV_DEV void load_tile(const double *A, int ld, int ti, int tj, double (*s)[CH_T + 1]) {
    for (int e = threadIdx.x; e < CH_T * CH_T; e += CH_THREADS) {
        int r = e >> 5, c = e & 31;
        s[r][c] = A[(size_t)(ti * CH_T + r) * ld + tj * CH_T + c];
    }
}
SDV_DEV void store_tile(double *A, int ld, int ti, int tj, double (*s)[CH_T + 1]) {
    for (int e = threadIdx.x; e < CH_T * CH_T; e += CH_THREADS) {
        int r = e >> 5, c = e & 31;
        A[(size_t)(ti * CH_T + r) * ld + tj * CH_T + c] = s[r][c];
    }
}

// in-place Cholesky of a 32x32 tile in shared memory by warp 0; invd receives 1/L_cc. Returns false if not PD.
SDV_DEV bool chol_tile_warp(double (*s)[CH_T + 1], double *invd, int lane) {
    bool ok = true;
    for (int c = 0; c < CH_T; c++) {
        double a0 = 0, a1 = 0;
        int q = 0;
        for (; q + 1 < c; q += 2) {
            a0 += s[lane][q] * s[c][q];
            a1 += s[lane][q + 1] * s[c][q + 1];
        }
        if (q < c) a0 += s[lane][q] * s[c][q];
        double v = s[lane][c] - (a0 + a1);
        double dcc = __shfl_sync(0xffffffffu, v, c);
        if (!(dcc > 0.0) || !isfinite(dcc)) {
            ok = false;
            dcc = 1.0;
        }
        double d = sqrt(dcc), id = 1.0 / d;
        if (lane == c) {
            s[c][c] = d;
            invd[c] = id;
        } else if (lane > c) {
            s[lane][c] = v * id;
        } else {
            s[lane][c] = 0.0; // strictly upper part of the tile is not part of L
        }
        __syncwarp();
    }
    return ok;
}

// X <- X * L^-T for a 32x32 tile (row r handled by lane r of the calling warp)
SDV_DEV void trsm_tile_warp(double (*x)[CH_T + 1], double (*L)[CH_T + 1], const double *invd, int lane) {
    for (int c = 0; c < CH_T; c++) {
        double a0 = 0, a1 = 0;
        int q = 0;
        for (; q + 1 < c; q += 2) {
            a0 += x[lane][q] * L[c][q];
            a1 += x[lane][q + 1] * L[c][q + 1];
        }
        if (q < c) a0 += x[lane][q] * L[c][q];
        x[lane][c] = (x[lane][c] - (a0 + a1)) * invd[c];
    }
}

__global__ void __launch_bounds__(CH_THREADS) k_chol_panel(double *A, double *Lo, int ld, int T, int k, const LMState *st, Accum *acc) {
    if (st->status != 0) return;
    __shared__ double sK[CH_T][CH_T + 1], sI[CH_T][CH_T + 1], sJ[CH_T][CH_T + 1];
    __shared__ double invd[CH_T];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npanel = T - k + 1;
    int ti, tj;
    bool panel = (int)blockIdx.x < npanel;
    if (panel) {
        ti = k + blockIdx.x;
        tj = k;
    } else {
        // decode trailing tile (i, j): rows i = k+1..T, cols j = k+1..min(i, T-1)
        int e = blockIdx.x - npanel;
        int i = k + 1;
        for (;;) {
            int cnt = min(i, T - 1) - k;
            if (e < cnt) break;
            e -= cnt;
            i++;
        }
        ti = i;
        tj = k + 1 + e;
    }
    load_tile(A, ld, k, k, sK);
    if (ti != k) load_tile(A, ld, ti, k, sI);
    if (!panel && tj != ti) load_tile(A, ld, tj, k, sJ);
    __syncthreads();
    if (warp == 0) {
        bool ok = chol_tile_warp(sK, invd, lane);
        if (!ok && lane == 0 && blockIdx.x == 0) acc->chol_fail = 1;
    }
    __syncthreads();
    if (panel) {
        if (ti == k) {
            store_tile(Lo, ld, k, k, sK);
        } else {
            if (warp == 0) trsm_tile_warp(sI, sK, invd, lane);
            __syncthreads();
            store_tile(Lo, ld, ti, k, sI);
        }
        return;
    }
    if (warp == 0) trsm_tile_warp(sI, sK, invd, lane);
    if (warp == 1 && tj != ti) trsm_tile_warp(sJ, sK, invd, lane);
    __syncthreads();
    double (*Bm)[CH_T + 1] = (tj == ti) ? sI : sJ;
    for (int e = threadIdx.x; e < CH_T * CH_T; e += CH_THREADS) {
        int r = e >> 5, c = e & 31;
        double s = 0;
#pragma unroll 8
        for (int q = 0; q < CH_T; q++) s += sI[r][q] * Bm[c][q];
        A[(size_t)(ti * CH_T + r) * ld + tj * CH_T + c] -= s;
    }
}

// =====================================================================================================================
// back substitution L^T z = y, delta_p = -z; candidate reduced parameters; model-decrease / norm terms of the reduced
// columns; frame-camera table of the candidate.  Single CTA.
// =====================================================================================================================
constexpr int TS_THREADS = 1024;
__global__ void __launch_bounds__(TS_THREADS) k_trisolve(const DevProblem *__restrict__ Pg, LinBuf B0, LinBuf B1, LMState *st, Accum *acc, const double *Lo,
                                                         const double *damp_p, const double *graw_p, double *dxp) {
    const DevProblem &P = *Pg; // device-resident problem description: the launch parameters do not depend on the window (one CUDA graph serves them all)
    if (st->status != 0) return;
    extern __shared__ double sh[];
    double *z = sh;                 // [n_pad]
    double *part = sh + P.n_pad;    // [32][32]
    double *dtile = part + 1024;    // [32][33]
    const int ld = P.ld, T = P.n_pad / CH_T, n = P.n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const LinBuf &Bx = st->cur ? B1 : B0;
    const LinBuf &Bc = st->cur ? B0 : B1;
    if (acc->chol_fail || acc->schur_fail) {
        if (threadIdx.x == 0) {
            st->step_valid = 0;
            st->model_cost_change = 0.0;
        }
        return;
    }
    for (int i = threadIdx.x; i < P.n_pad; i += blockDim.x) z[i] = Lo[(size_t)P.n_pad * ld + i]; // y = L^-1 g
    __syncthreads();
    for (int kt = T - 1; kt >= 0; kt--) {
        // diagonal tile -> shared memory (one element per thread)
        {
            int r = threadIdx.x >> 5, c = threadIdx.x & 31;
            dtile[r * 33 + c] = Lo[(size_t)(kt * CH_T + r) * ld + kt * CH_T + c];
        }
        // s_c = sum_{i>kt} sum_r L[i*32+r][kt*32+c] z[i*32+r]
        double s = 0;
        for (int it = kt + 1 + warp; it < T; it += 32) {
            const double *Lt = Lo + (size_t)(it * CH_T) * ld + kt * CH_T + lane;
#pragma unroll 8
            for (int r = 0; r < CH_T; r++) s += Lt[(size_t)r * ld] * z[it * CH_T + r];
        }
        part[warp * 32 + lane] = s;
        __syncthreads();
        if (warp == 0) {
            double tot = 0;
            for (int w2 = 0; w2 < 32; w2++) tot += part[w2 * 32 + lane];
            double y = z[kt * CH_T + lane] - tot;
            // solve L_kk^T x = y : x_c = (y_c - sum_{r>c} L[r][c] x_r) / L[c][c], c = 31 .. 0
            double x = 0;
            for (int c = CH_T - 1; c >= 0; c--) {
                double xc = __shfl_sync(0xffffffffu, y, c) / dtile[c * 33 + c];
                if (lane == c) x = xc;
                if (lane < c) y -= dtile[c * 33 + lane] * xc; // equation `lane` contains L[c][lane] x_c
            }
            z[kt * CH_T + lane] = x;
        }
        __syncthreads();
    }
    // delta_p = -z ; candidate = x + delta ; reductions
    double gd = 0, dd = 0, sn = 0, cn = 0;
    for (int i = threadIdx.x; i < P.n_pad; i += blockDim.x) {
        double d = i < n ? -z[i] : 0.0;
        dxp[i] = d;
        double xc = Bx.xp[i] + d;
        Bc.xp[i] = i < n ? xc : 0.0;
        if (i < n) {
            gd += graw_p[i] * d;
            dd += damp_p[i] * d * d;
            sn += d * d;
            cn += xc * xc;
        }
    }
    gd = warp_sum(gd);
    dd = warp_sum(dd);
    sn = warp_sum(sn);
    cn = warp_sum(cn);
    if (lane == 0 && P.rank == 0) { // every rank holds the same reduced system: count it once
        atomicAdd(&acc->model_gd, gd);
        atomicAdd(&acc->model_dd, dd);
        atomicAdd(&acc->step_norm2, sn);
        atomicAdd(&acc->cand_norm2, cn);
    }
    __syncthreads();
    // frame-camera table of the candidate point
    for (int i = threadIdx.x; i < P.F * P.C; i += blockDim.x) compute_fct_row(P, Bc.xp, i / P.C, i % P.C, Bc.fct + (size_t)i * FCT_ROW);
    if (threadIdx.x == 0) st->step_valid = 1;
}

// =====================================================================================================================
// LM control (Ceres 2.2 TrustRegionMinimizer + LevenbergMarquardtStrategy), single thread
// =====================================================================================================================
__global__ void k_ctrl_init(const DevProblem *__restrict__ Pg, LMState *st, Accum *acc, SolverOpts opt) {
    // after the iteration-0 linearisation into buffer 0
    st->iter = 0;
    st->status = 0;
    st->cur = 0;
    st->step_valid = 0;
    st->have_cand = 0;
    st->num_consecutive_invalid = 0;
    st->atleast_one_successful = 0;
    st->n_ok = st->n_bad = 0;
    st->scaling_done = 0;
    st->need_grad_check = 1;
    st->radius = opt.initial_radius;
    st->decrease_factor = 2.0;
    st->x_cost = acc->cost[0];
    st->initial_cost = acc->cost[0];
    st->cand_cost = 0.0;
    st->model_cost_change = 0.0;
    st->x_norm2 = 0.0;
    st->trace_cost[0] = acc->cost[0];
    st->trace_radius[0] = opt.initial_radius;
    st->trace_model[0] = 0.0;
    st->trace_accepted[0] = 1;
    acc->cost[1] = 0.0;
    acc->model_gd = acc->model_dd = acc->step_norm2 = acc->cand_norm2 = 0.0;
    acc->grad_max_bits = 0ull;
    acc->schur_fail = acc->chol_fail = 0;
    if ((Pg->max_iter > 0 ? Pg->max_iter : opt.max_num_iterations) <= 0) st->status = 1 + 0;
    else st->iter = 1; // Ceres bumps its iteration counter before computing the step (was a kernel of its own)
}

SDV_DEV void ctrl_step(LMState *st, Accum *acc, const SolverOpts &opt, int max_iter);
// One warp: the solver state (1.8 KB) and the accumulators are staged through shared memory with coalesced loads / stores, the
// control logic itself runs on lane 0 — a single thread walking the structures in global memory took 5.5 us per iteration.
__global__ void __launch_bounds__(32) k_ctrl(const DevProblem *__restrict__ Pg, LMState *st_g, Accum *acc_g, SolverOpts opt, unsigned long long cond) {
    // `cond` (0 = none) is the handle of the CUDA-graph WHILE node whose body is one LM iteration
    __shared__ LMState st_s;
    __shared__ Accum acc_s;
    static_assert(sizeof(LMState) % 8 == 0 && sizeof(Accum) % 8 == 0, "staged as 64-bit words");
    const int lane = threadIdx.x;
    if (st_g->status != 0) {
        if (lane == 0 && cond) cudaGraphSetConditional((cudaGraphConditionalHandle)cond, 0);
        return;
    }
    {
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(st_g);
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(&st_s);
        for (int i = lane; i < (int)(sizeof(LMState) / 8); i += 32) dst[i] = src[i];
        src = reinterpret_cast<const unsigned long long *>(acc_g);
        dst = reinterpret_cast<unsigned long long *>(&acc_s);
        for (int i = lane; i < (int)(sizeof(Accum) / 8); i += 32) dst[i] = src[i];
    }
    __syncwarp();
    if (lane == 0) {
        ctrl_step(&st_s, &acc_s, opt, Pg->max_iter > 0 ? Pg->max_iter : opt.max_num_iterations); // per-window cap: 10 / 5 in the f2 solves (AOptimizer.cpp:113,166,247)
        if (cond) cudaGraphSetConditional((cudaGraphConditionalHandle)cond, st_s.status == 0 ? 1u : 0u);
    }
    __syncwarp();
    {
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(&st_s);
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(st_g);
        for (int i = lane; i < (int)(sizeof(LMState) / 8); i += 32) dst[i] = src[i];
        src = reinterpret_cast<const unsigned long long *>(&acc_s);
        dst = reinterpret_cast<unsigned long long *>(acc_g);
        for (int i = lane; i < (int)(sizeof(Accum) / 8); i += 32) dst[i] = src[i];
    }
}

SDV_DEV void ctrl_step(LMState *st, Accum *acc, const SolverOpts &opt, int max_iter) {
    const int it = st->iter;
    const int ti = it < 63 ? it : 63;
    const int cand = 1 - st->cur;
    double model_cost_change = 0.0;
    bool valid = st->step_valid != 0;
    if (valid) {
        // -(J step)^T (r + J step / 2) = -1/2 g.delta + 1/2 delta^T D delta   (with (J^T J + D) delta = -g)
        model_cost_change = -0.5 * acc->model_gd + 0.5 * acc->model_dd;
        valid = model_cost_change > 0.0;
    }
    st->trace_model[ti] = model_cost_change;
    st->model_cost_change = model_cost_change;
    if (!valid) {
        // HandleInvalidStep
        st->num_consecutive_invalid += 1;
        st->n_bad += 1;
        st->trace_cost[ti] = st->x_cost;
        st->trace_accepted[ti] = -1;
        if (st->num_consecutive_invalid >= opt.max_consecutive_invalid_steps) {
            st->status = 1 + 5;
            st->trace_radius[ti] = st->radius;
        } else {
            st->radius *= 0.5;
            st->trace_radius[ti] = st->radius;
        }
    } else {
        st->num_consecutive_invalid = 0;
        const double cand_cost = acc->cost[cand];
        st->cand_cost = cand_cost;
        const double step_norm = sqrt(acc->step_norm2);
        const double x_norm = sqrt(st->x_norm2);
        bool stop = false;
        if (st->atleast_one_successful && step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) {
            st->status = 1 + 3;
            stop = true;
        }
        if (!stop && fabs(st->x_cost - cand_cost) <= opt.function_tolerance * st->x_cost) {
            st->status = 1 + 1; // candidate NOT applied
            stop = true;
        }
        if (stop) {
            st->trace_cost[ti] = st->x_cost;
            st->trace_radius[ti] = st->radius;
            st->trace_accepted[ti] = 0;
        } else {
            const double rho = (st->x_cost - cand_cost) / model_cost_change;
            if (rho > opt.min_relative_decrease) {
                st->cur = cand; // x <- candidate (its linearisation is already in that buffer)
                st->x_cost = cand_cost;
                st->x_norm2 = acc->cand_norm2;
                double t = 2.0 * rho - 1.0;
                st->radius = st->radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
                st->radius = fmin(opt.max_radius, st->radius);
                st->decrease_factor = 2.0;
                st->atleast_one_successful = 1;
                st->need_grad_check = 1;
                st->n_ok += 1;
                st->trace_accepted[ti] = 1;
            } else {
                st->radius = st->radius / st->decrease_factor;
                st->decrease_factor *= 2.0;
                st->n_bad += 1;
                st->trace_accepted[ti] = 0;
            }
            st->trace_cost[ti] = st->x_cost;
            st->trace_radius[ti] = st->radius;
        }
    }
    // FinalizeIterationAndCheckIfMinimizerCanContinue for the next trip
    if (st->status == 0) {
        if (it >= max_iter) st->status = 1 + 0;
        else if (st->radius <= opt.min_radius) st->status = 1 + 4;
    }
    // reset accumulators for the next iteration
    acc->cost[1 - st->cur] = 0.0;
    acc->model_gd = acc->model_dd = acc->step_norm2 = acc->cand_norm2 = 0.0;
    acc->schur_fail = acc->chol_fail = 0;
    if (st->status == 0) { // start of the next iteration
        st->iter += 1;
        st->step_valid = 0;
    }
}

// start of a solve: x = 0 in both linearisation buffers, solver state and accumulators cleared (one launch instead of six
// memset nodes whose sizes would tie the CUDA graph to the window)
// (Sz / nz: the reduced-system buffer, zeroed here for the first iteration; k_backsub_cost zeroes it for the following ones)
__global__ void k_reset(const DevProblem *__restrict__ Pg, LinBuf B0, LinBuf B1, LMState *st, Accum *acc, double *Sz = nullptr, long long nz = 0) {
    const DevProblem &P = *Pg;
    for (long long i = 2 * ((long long)blockIdx.x * blockDim.x + threadIdx.x); i < nz; i += 2 * (long long)gridDim.x * blockDim.x)
        *reinterpret_cast<double2 *>(Sz + i) = make_double2(0.0, 0.0);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int i = tid; i < P.n_pad; i += nt) {
        B0.xp[i] = 0.0;
        B1.xp[i] = 0.0;
    }
    const int n3 = 3 * (P.L > 0 ? P.L : 1);
    for (int i = tid; i < n3; i += nt) {
        B0.xl[i] = 0.0;
        B1.xl[i] = 0.0;
    }
    static_assert(sizeof(LMState) % 8 == 0 && sizeof(Accum) % 8 == 0, "cleared as 64-bit words");
    unsigned long long *w = reinterpret_cast<unsigned long long *>(st);
    for (int i = tid; i < (int)(sizeof(LMState) / 8); i += nt) w[i] = 0ull;
    w = reinterpret_cast<unsigned long long *>(acc);
    for (int i = tid; i < (int)(sizeof(Accum) / 8); i += nt) w[i] = 0ull;
}

// gather the solution blocks in ABI order, followed by the solver state and the accumulators:
// out = [dpose 6F | dv 3F | dba 3F | dbg 3F | LMState | Accum | dlmk 3 max(L,1)] — ONE device-to-host copy per solve; the landmark
// block comes last so that the multi-GPU gather (a sum over ranks of that block) can run over its CAPACITY
__global__ void k_gather_solution(const DevProblem *__restrict__ Pg, LinBuf B0, LinBuf B1, const LMState *st, const Accum *acc, double *out) {
    const DevProblem &P = *Pg;
    const LinBuf &B = st->cur ? B1 : B0;
    double *dpose = out, *dv = out + 6 * P.F, *dba = dv + 3 * P.F, *dbg = dba + 3 * P.F;
    unsigned long long *tail = reinterpret_cast<unsigned long long *>(dbg + 3 * P.F);
    double *dlmk = reinterpret_cast<double *>(tail + (sizeof(LMState) + sizeof(Accum)) / 8);
    int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int f = tid; f < P.F; f += nt) {
        int pc = P.pose_col[f], vc = P.vb_col[f];
        for (int k = 0; k < 6; k++) dpose[6 * f + k] = pc >= 0 ? B.xp[pc + k] : 0.0;
        for (int k = 0; k < 3; k++) {
            dv[3 * f + k] = vc >= 0 ? B.xp[vc + k] : 0.0;
            dba[3 * f + k] = vc >= 0 ? B.xp[vc + 3 + k] : 0.0;
            dbg[3 * f + k] = vc >= 0 ? B.xp[vc + 6 + k] : 0.0;
        }
    }
    for (int l = tid; l < P.L; l += nt) {
        int dc = P.lmk_col[l];
        // kept landmarks are replicated on every rank (count them once); eliminated ones are owned by one rank
        for (int k = 0; k < 3; k++) dlmk[3 * (size_t)l + k] = dc >= 0 ? (P.rank == 0 ? B.xp[dc + k] : 0.0) : B.xl[3 * (size_t)l + k];
    }
    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(st);
    for (int i = tid; i < (int)(sizeof(LMState) / 8); i += nt) tail[i] = src[i];
    tail += sizeof(LMState) / 8;
    src = reinterpret_cast<const unsigned long long *>(acc);
    for (int i = tid; i < (int)(sizeof(Accum) / 8); i += nt) tail[i] = src[i];
}

} // namespace sdv

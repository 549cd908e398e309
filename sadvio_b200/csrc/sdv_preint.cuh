// IMU pre-integration on the device: isae::IMU::processIMU (reference cpp/src/data/sensors/IMU.cpp:5-91) chained over the
// samples of every keyframe interval of a window, one warp per interval.  It produces what IMUFactor (residuals.hpp:133-245)
// consumes — DeltaR, DeltaV, DeltaP, the 9x9 covariance and the five bias Jacobians — plus the dead-reckoned pose / velocity
// the same call leaves in the frame (IMU.cpp:34-41).
//
// The recursion is sequential in the samples (~50 per interval) and independent across intervals.  All lanes carry the 3x3
// quantities redundantly in registers; the covariance propagation Sigma' = A Sigma A^T + B eta B^T (two 9x9x9 products per
// sample) is spread over the warp, one matrix entry per lane and pass, through shared memory.
//
// Reference quirks kept: dt > 1 s is replaced by 1 / rate (IMU.cpp:23-25); the noise matrix B of the FIRST step after a
// keyframe reads the keyframe IMU's own, stale, _delta_R (IMU.cpp:44-47 precede the restart at :50).
#pragma once
#include "sdv_kernels.cuh"

namespace sdv {

struct PreintArgs {
    int n_intervals;
    const int *sample_ptr;                 // [n+1]
    const double *acc, *gyr, *dt;          // [S][3], [S][3], [S]
    const double *T_f_w, *v, *ba, *bg;     // keyframe state at the start of each interval: [n][12], [n][3] x 3
    const double *dR_stale;                // [n][9] or nullptr (identity)
    double eta[6], rate_hz;
    // outputs, [n][...]
    double *dR, *dv, *dp, *cov, *J_dR_bg, *J_dv_ba, *J_dv_bg, *J_dp_ba, *J_dp_bg, *T_pred, *v_pred;
};

constexpr int PRE_WARPS = 4;

__global__ void __launch_bounds__(PRE_WARPS * 32) k_preintegrate(PreintArgs a) {
    __shared__ double Sg[PRE_WARPS][81], Tm[PRE_WARPS][81], Am[PRE_WARPS][81], Qm[PRE_WARPS][81];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int iv = blockIdx.x * PRE_WARPS + wib;
    if (iv >= a.n_intervals) return;
    double *S = Sg[wib], *T = Tm[wib], *A = Am[wib], *Q = Qm[wib];
    const double g[3] = {0.0, 0.0, -9.81}; // IMU.h:8
    // keyframe state
    double Rfw[9], tfw[3], Rwf[9], twf[3], v[3], ba[3], bg[3], lastR[9];
    load_RT(a.T_f_w + 12 * (size_t)iv, Rfw, tfw);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) Rwf[i * 3 + j] = Rfw[j * 3 + i];
    matT3_vec(Rfw, tfw, twf);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        twf[k] = -twf[k];
        v[k] = a.v[3 * (size_t)iv + k];
        ba[k] = a.ba[3 * (size_t)iv + k];
        bg[k] = a.bg[3 * (size_t)iv + k];
    }
#pragma unroll
    for (int k = 0; k < 9; k++) lastR[k] = a.dR_stale ? a.dR_stale[9 * (size_t)iv + k] : (k % 4 == 0 ? 1.0 : 0.0);
    double dRs[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, dvs[3] = {0, 0, 0}, dps[3] = {0, 0, 0};
    double JRg[9], Jva[9], Jvg[9], Jpa[9], Jpg[9];
#pragma unroll
    for (int k = 0; k < 9; k++) JRg[k] = Jva[k] = Jvg[k] = Jpa[k] = Jpg[k] = 0.0;
    const int s0 = a.sample_ptr[iv], s1 = a.sample_ptr[iv + 1];
    for (int s = s0; s < s1; s++) {
        double dt = a.dt[s];
        if (dt > 1) dt = 1 / a.rate_hz;                                       // IMU.cpp:23-25
        const double dt22 = 0.5 * dt * dt;
        double am[3], wm[3], dv[3], dp[3], dR[9], Jrk[9];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            am[k] = a.acc[3 * (size_t)s + k] - ba[k];
            wm[k] = (a.gyr[3 * (size_t)s + k] - bg[k]) * dt;
            dv[k] = am[k] * dt;                                                 // :28
            dp[k] = am[k] * dt22;                                               // :29
        }
        exp_so3(wm, dR);                                                        // :30
        right_jacobian(wm, Jrk);                                                // :31 (the keyframe's bg: the same value)
        // velocity / pose prediction (:34-41), with the OLD rotation and velocity
        double Rdv[3], Rdp[3], Rn[9];
        mat3_vec(Rwf, dv, Rdv);
        mat3_vec(Rwf, dp, Rdp);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            twf[k] += v[k] * dt + Rdp[k] + g[k] * dt22;
            v[k] = v[k] + Rdv[k] + g[k] * dt;
        }
        mat3_mul(Rwf, dR, Rn);
#pragma unroll
        for (int k = 0; k < 9; k++) Rwf[k] = Rn[k];
        // Q = B eta B^T (+ integration covariance), B = [Jrk dt 0; 0 lastR dt; 0 lastR dt22]   (:44-47, :54-55 / :80-81)
        for (int e = lane; e < 81; e += 32) {
            const int i = e / 9, j = e - 9 * i, bi = i / 3, bj = j / 3, ii = i - 3 * bi, jj = j - 3 * bj;
            double q = 0.0;
            if (bi == 0 && bj == 0) {
#pragma unroll
                for (int k = 0; k < 3; k++) q += (Jrk[ii * 3 + k] * dt) * a.eta[k] * (Jrk[jj * 3 + k] * dt);
            } else if (bi >= 1 && bj >= 1) {
                const double si = bi == 1 ? dt : dt22, sj = bj == 1 ? dt : dt22;
#pragma unroll
                for (int k = 0; k < 3; k++) q += (lastR[ii * 3 + k] * si) * a.eta[3 + k] * (lastR[jj * 3 + k] * sj);
            }
            if (i == j && i >= 6) q += 0.0001 * dt;
            Q[e] = q;
        }
        if (s == s0) { // restart of the pre-integration: the previous measurement sits in a keyframe (:50-61)
#pragma unroll
            for (int k = 0; k < 9; k++) {
                dRs[k] = dR[k];
                JRg[k] = -(Jrk[k] * dt);
                Jva[k] = k % 4 == 0 ? -dt : 0.0;
                Jvg[k] = 0.0;
                Jpa[k] = k % 4 == 0 ? -dt22 : 0.0;
                Jpg[k] = 0.0;
            }
#pragma unroll
            for (int k = 0; k < 3; k++) {
                dvs[k] = dv[k];
                dps[k] = dp[k];
            }
            __syncwarp();
            for (int e = lane; e < 81; e += 32) S[e] = Q[e];
            __syncwarp();
        } else { // :63-88
            double dRdA[9], Sk[9], t3[3], nR[9];
            skew3(am, Sk);
            mat3_mul(dRs, Sk, dRdA);                                            // last DeltaR * [acc - ba]x
            for (int e = lane; e < 81; e += 32) {
                const int i = e / 9, j = e - 9 * i, bi = i / 3, bj = j / 3, ii = i - 3 * bi, jj = j - 3 * bj;
                double x = i == j ? 1.0 : 0.0;
                if (bi == 0 && bj == 0) x = dR[jj * 3 + ii];                    // dR^T
                else if (bi == 1 && bj == 0) x = -(dRdA[ii * 3 + jj] * dt);
                else if (bi == 2 && bj == 0) x = -(dRdA[ii * 3 + jj] * dt22);
                else if (bi == 2 && bj == 1) x = ii == jj ? dt : 0.0;
                A[e] = x;
            }
            __syncwarp();
            for (int e = lane; e < 81; e += 32) { // T = A Sigma
                const int i = e / 9, j = e - 9 * i;
                double x = 0.0;
#pragma unroll
                for (int k = 0; k < 9; k++) x += A[i * 9 + k] * S[k * 9 + j];
                T[e] = x;
            }
            __syncwarp();
            for (int e = lane; e < 81; e += 32) { // Sigma' = T A^T + Q
                const int i = e / 9, j = e - 9 * i;
                double x = 0.0;
#pragma unroll
                for (int k = 0; k < 9; k++) x += T[i * 9 + k] * A[j * 9 + k];
                S[e] = x + Q[e];
            }
            __syncwarp();
            // bias Jacobians with the LAST deltas (:84-88), then the deltas themselves (:66-68)
            double M1[9], M2[9], nJRg[9], dRT[9];
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) dRT[i * 3 + j] = dR[j * 3 + i];
            mat3_mul(dRT, JRg, nJRg);
            mat3_mul(dRdA, JRg, M1); // dR_dA * last J_dR_bg
#pragma unroll
            for (int k = 0; k < 9; k++) {
                const double jva = Jva[k], jvg = Jvg[k];
                Jpa[k] = Jpa[k] + jva * dt - dt22 * dRs[k];
                Jpg[k] = Jpg[k] + jvg * dt - dt22 * M1[k];
                Jva[k] = jva - dRs[k] * dt;
                Jvg[k] = jvg - M1[k] * dt;
                JRg[k] = nJRg[k] - Jrk[k] * dt;
            }
            (void)M2;
            mat3_vec(dRs, dp, t3);
#pragma unroll
            for (int k = 0; k < 3; k++) dps[k] = dps[k] + dvs[k] * dt + t3[k];
            mat3_vec(dRs, dv, t3);
#pragma unroll
            for (int k = 0; k < 3; k++) dvs[k] = dvs[k] + t3[k];
            mat3_mul(dRs, dR, nR);
#pragma unroll
            for (int k = 0; k < 9; k++) dRs[k] = nR[k];
        }
#pragma unroll
        for (int k = 0; k < 9; k++) lastR[k] = dRs[k];
    }
    // outputs
    const size_t o = (size_t)iv;
    if (lane < 9) {
        a.dR[9 * o + lane] = dRs[lane];
        a.J_dR_bg[9 * o + lane] = JRg[lane];
        a.J_dv_ba[9 * o + lane] = Jva[lane];
        a.J_dv_bg[9 * o + lane] = Jvg[lane];
        a.J_dp_ba[9 * o + lane] = Jpa[lane];
        a.J_dp_bg[9 * o + lane] = Jpg[lane];
    }
    if (lane < 3) {
        a.dv[3 * o + lane] = dvs[lane];
        a.dp[3 * o + lane] = dps[lane];
        if (a.v_pred) a.v_pred[3 * o + lane] = v[lane];
    }
    __syncwarp();
    for (int e = lane; e < 81; e += 32) a.cov[81 * o + e] = s1 > s0 ? S[e] : 0.0;
    if (a.T_pred && lane == 0) { // T_f_w = (T_w_f)^-1
        double Rt[9], t[3];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) Rt[i * 3 + j] = Rwf[j * 3 + i];
        mat3_vec(Rt, twf, t);
        double *Tp = a.T_pred + 12 * o;
        for (int i = 0; i < 3; i++) {
            for (int j = 0; j < 3; j++) Tp[i * 4 + j] = Rt[i * 3 + j];
            Tp[i * 4 + 3] = -t[i];
        }
    }
}

} // namespace sdv

// TEST INFRASTRUCTURE — CPU oracle (see oracle/README.md). Never imported by the product path.
//
// Behavioural restatement of the reference's SO(3) helpers and hot-path cost functors.
// Every function cites the reference file:line (relative to /root/reference/cpp) it follows.
// Evaluate() signatures follow ceres::CostFunction::Evaluate: parameters[k] is the k-th parameter
// block, jacobians[k] (may be NULL) receives a ROW-MAJOR num_residuals x block_size matrix.
#pragma once
#include "smallmat.hpp"
#include <algorithm>
#include <cmath>

namespace orc {

// ---------------------------------------------------------------------------------------------
// include/utilities/geometry.h
// ---------------------------------------------------------------------------------------------
inline M3 skewMatrix(const V3 &w) { // geometry.h:17-23
    M3 s;
    s(0, 0) = 0;      s(0, 1) = -w[2];  s(0, 2) = w[1];
    s(1, 0) = w[2];   s(1, 1) = 0;      s(1, 2) = -w[0];
    s(2, 0) = -w[1];  s(2, 1) = w[0];   s(2, 2) = 0;
    return s;
}
inline V3 FromskewMatrix(const M3 &s) { return vec3(s(2, 1), s(0, 2), s(1, 0)); } // geometry.h:25-28

inline M3 so3_rightJacobian(const V3 &w) { // geometry.h:30-37 (identity below 1e-5)
    double w_norm = w.norm();
    M3 w_skew = skewMatrix(w);
    if (w_norm < 1e-5) return M3::Identity();
    return M3::Identity() - ((1 - std::cos(w_norm)) / (w_norm * w_norm)) * w_skew +
           ((w_norm - std::sin(w_norm)) / (w_norm * w_norm * w_norm)) * (w_skew * w_skew);
}

inline M3 exp_so3(const V3 &v) { // geometry.h:131-147 (first order below 1e-9)
    double tolerance = 1e-9;
    double angle = v.norm();
    M3 Rot;
    if (angle < tolerance) {
        Rot = M3::Identity() + skewMatrix(v);
    } else {
        V3 axis = v / angle;
        M3 skew = skewMatrix(axis);
        Rot = M3::Identity() + (1. - std::cos(angle)) * (skew * skew) + std::sin(angle) * skew;
    }
    return Rot;
}

inline V3 log_so3(const M3 &M) { // geometry.h:149-166
    double tolerance = 1e-9;
    double cos_angle = 0.5 * M.trace() - 0.5;
    cos_angle = std::min(std::max(cos_angle, -1.), 1.);
    double angle = std::acos(cos_angle);
    V3 phi;
    if (std::fabs(std::sin(angle)) < tolerance || angle < tolerance)
        phi = 0.5 * FromskewMatrix(M - M.T());
    else
        phi = (0.5 * angle / std::sin(angle)) * FromskewMatrix(M - M.T());
    return phi;
}

inline Aff se3_doubleVec6dtoRT(const double *pose) { // geometry.h:202-207
    Aff RT;
    RT.R = exp_so3(vec3(pose[0], pose[1], pose[2]));
    RT.t = vec3(pose[3], pose[4], pose[5]);
    return RT;
}
inline Mat<6, 1> se3_RTtoVec6d(const Aff &RT) { // geometry.h:168-178
    Mat<6, 1> pose;
    V3 w = log_so3(RT.R);
    for (int i = 0; i < 3; i++) {
        pose[i] = w[i];
        pose[3 + i] = RT.t[i];
    }
    return pose;
}

static const double kGravity[3] = {0, 0, -9.81}; // data/sensors/IMU.h:8

// ---------------------------------------------------------------------------------------------
// AngularErrCeres_pointxd_dx  — AngularAdjustmentCERESAnalytic.h:45-120, SizedCostFunction<2,6,3>
// ---------------------------------------------------------------------------------------------
struct AngularErr {
    V3 bearing;
    Aff T_s_f, T_f_w;
    V3 t_w_lmk;
    double sigma;
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const {
        Aff dT = se3_doubleVec6dtoRT(parameters[0]);                    // .h:57
        V3 dt = V3::From(parameters[1]);                                // .h:58
        double weight = 1 / sigma;                                      // .h:59
        V3 t_s_lmk = T_s_f * (T_f_w * (dT * (t_w_lmk + dt)));           // .h:62
        double t_s_lmk_norm = t_s_lmk.norm();
        V3 b_s_lmk = t_s_lmk / t_s_lmk_norm;                            // .h:64
        V3 b1;                                                          // .h:67-74
        if ((bearing - vec3(1, 0, 0)).norm() > 1e-5) {
            b1 = cross(bearing, vec3(1, 0, 0));
            b1 = b1 / b1.norm();
        } else {
            b1 = cross(bearing, vec3(0, 0, 1));
            b1 = b1 / b1.norm();
        }
        V3 b2 = cross(b1, bearing);                                     // .h:76-77
        b2 = b2 / b2.norm();
        Mat<3, 2> P;
        for (int i = 0; i < 3; i++) {
            P(i, 0) = b1[i];
            P(i, 1) = b2[i];
        }
        Mat<2, 3> Pt = P.T();
        Mat<2, 1> res = weight * (Pt * (b_s_lmk - bearing));            // .h:85
        residuals[0] = res[0];
        residuals[1] = res[1];
        if (jacobians != nullptr) {
            Mat<2, 3> J_e_lmk = (Pt * (M3::Identity() - b_s_lmk * b_s_lmk.T()) * T_s_f.R * T_f_w.R) / t_s_lmk_norm; // .h:89-91
            if (jacobians[0] != nullptr) {
                Mat<3, 6> J_bear_frame = Mat<3, 6>::Zero();             // .h:94-99
                V3 w_dT = se3_RTtoVec6d(dT).block<3, 1>(0, 0);
                J_bear_frame.setBlock(0, 0, -(dT.R * skewMatrix(t_w_lmk + dt) * so3_rightJacobian(w_dT)));
                J_bear_frame.setBlock(0, 3, M3::Identity());
                Mat<2, 6> J_frame = weight * (J_e_lmk * J_bear_frame);  // .h:101
                J_frame.to(jacobians[0]);
            }
            if (jacobians[1] != nullptr) {
                Mat<2, 3> J_lmk = weight * (J_e_lmk * dT.R);            // .h:105
                J_lmk.to(jacobians[1]);
            }
        }
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// Camera::project with Jacobians — data/sensors/Camera.cpp:84-139
// K = (fx, fy, cx, cy); sqrt_info 2x2.
// ---------------------------------------------------------------------------------------------
inline bool camera_project(const double K[4], const Aff &T_s_f, const V3 &t_w_lmk_full, const Aff &T_f_w,
                           const Mat<2, 2> &sqrt_info, Mat<2, 1> &p2d, double *J_proj_frame, double *J_proj_lmk) {
    M3 Kmat = M3::Identity();
    Kmat(0, 0) = K[0];
    Kmat(1, 1) = K[1];
    Kmat(0, 2) = K[2];
    Kmat(1, 2) = K[3];
    V3 t_cam_lmk = T_s_f * (T_f_w * t_w_lmk_full);                      // Camera.cpp:92-93
    V3 pt = Kmat * t_cam_lmk;                                           // :96
    Mat<2, 3> J_h;                                                      // :97-99
    J_h(0, 0) = 1 / pt[2];  J_h(0, 1) = 0.0;        J_h(0, 2) = -pt[0] / (pt[2] * pt[2]);
    J_h(1, 0) = 0.0;        J_h(1, 1) = 1 / pt[2];  J_h(1, 2) = -pt[1] / (pt[2] * pt[2]);
    pt = pt / pt[2];                                                    // :101
    p2d[0] = pt[0];
    p2d[1] = pt[1];
    if (J_proj_frame != nullptr) {                                      // :104-117
        Mat<3, 6> J_int = Mat<3, 6>::Zero();
        V3 w = se3_RTtoVec6d(T_f_w).block<3, 1>(0, 0);
        J_int.setBlock(0, 0, -(T_f_w.R * skewMatrix(t_w_lmk_full) * so3_rightJacobian(w)));
        J_int.setBlock(0, 3, M3::Identity());
        Mat<2, 6> J_frame = J_h * Kmat * T_s_f.R * J_int;
        J_frame = sqrt_info * J_frame;
        J_frame.to(J_proj_frame);
    }
    if (J_proj_lmk != nullptr) {                                        // :119-127
        M3 J_aug = T_s_f.R * T_f_w.R;
        Mat<2, 3> J_lmk = J_h * Kmat * J_aug;
        J_lmk = sqrt_info * J_lmk;
        J_lmk.to(J_proj_lmk);
    }
    if (t_cam_lmk[2] < 0.1) return false;                               // :128-129
    if (p2d[0] < 0 || p2d[1] < 0 || p2d[0] > 2 * K[2] || p2d[1] > 2 * K[3]) return false; // :131-133
    if (!std::isfinite(p2d[0]) || !std::isfinite(p2d[1])) return false; // :135-136
    return true;
}

// ---------------------------------------------------------------------------------------------
// ReprojectionErrCeres_pointxd_dx — BundleAdjustmentCERESAnalytic.h:41-98, SizedCostFunction<2,6,3>
// ---------------------------------------------------------------------------------------------
struct ReprojErr {
    Mat<2, 1> p2d;
    double K[4];
    Aff T_s_f, T_f_w_base; // cam_->getFrame()->getWorld2FrameTransform()
    V3 t_w_lmk;            // _T_w_lmk translation (point landmarks have identity orientation,
                           // landmarkinitializer/Point3DlandmarkInitializer.cpp:83-93)
    double sigma;
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const {
        Mat<2, 2> info_sqrt = (1 / sigma) * Mat<2, 2>::Identity();      // .h:49
        Aff T_f_w = T_f_w_base * se3_doubleVec6dtoRT(parameters[0]);    // .h:54-55
        V3 t_full = t_w_lmk + V3::From(parameters[1]);                  // .h:58
        Mat<2, 1> projection;
        if (jacobians != nullptr) {
            bool ok = camera_project(K, T_s_f, t_full, T_f_w, info_sqrt, projection, jacobians[0], jacobians[1]);
            Mat<2, 1> res = ok ? info_sqrt * (projection - p2d) : Mat<2, 1>::Zero(); // .h:63-68 (J kept on failure)
            residuals[0] = res[0];
            residuals[1] = res[1];
            if (jacobians[0] != nullptr) {                              // .h:70-80
                Mat<2, 6> J_proj_f = Mat<2, 6>::From(jacobians[0]);
                Mat<6, 6> J_lf_dlf = Mat<6, 6>::Zero();
                V3 dw = vec3(parameters[0][0], parameters[0][1], parameters[0][2]);
                J_lf_dlf.setBlock(0, 0, inverse3(so3_rightJacobian(log_so3(T_f_w.R))) * so3_rightJacobian(dw));
                J_lf_dlf.setBlock(3, 3, T_f_w_base.R);
                J_proj_f = J_proj_f * J_lf_dlf;
                J_proj_f.to(jacobians[0]);
            }
        } else {
            bool ok = camera_project(K, T_s_f, t_full, T_f_w, info_sqrt, projection, nullptr, nullptr); // .h:82-88
            Mat<2, 1> res = ok ? info_sqrt * (projection - p2d) : Mat<2, 1>::Zero();
            residuals[0] = res[0];
            residuals[1] = res[1];
        }
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// IMUFactor — residuals.hpp:133-245, SizedCostFunction<9, 6,6,3,3,3,3>
// parameters: dT_i, dT_j, dv_i, dv_j, dba_i, dbg_i
// ---------------------------------------------------------------------------------------------
struct ImuFactor {
    Aff T_fi_w_base, T_fj_w_base;
    V3 v_i_base, v_j_base;
    double dtij;
    Mat<9, 9> cov;          // _imu_j->getCov()
    M3 delta_R;             // getDeltaR()
    V3 delta_v, delta_p;    // getDeltaV(), getDeltaP()
    M3 J_dR_bg, J_dv_ba, J_dv_bg, J_dp_ba, J_dp_bg;
    static bool InfSqrt(const Mat<9, 9> &cov, Mat<9, 9> &inf_sqrt) {   // residuals.hpp:151-154
        Mat<9, 9> inv, L;
        if (!inverseLU<9>(cov, inv)) return false;
        if (!choleskyL<9>(inv, L)) return false;
        inf_sqrt = L.T();
        return true;
    }
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const {
        Aff T_fi_w = T_fi_w_base * se3_doubleVec6dtoRT(parameters[0]);  // :140-143
        Aff T_fj_w = T_fj_w_base * se3_doubleVec6dtoRT(parameters[1]);
        V3 v_i = v_i_base + V3::From(parameters[2]);                    // :144-145
        V3 v_j = v_j_base + V3::From(parameters[3]);
        V3 d_ba = V3::From(parameters[4]);
        V3 d_bg = V3::From(parameters[5]);
        V3 g = V3::From(kGravity);
        Mat<9, 9> inf_sqrt;
        if (!InfSqrt(cov, inf_sqrt)) return false;
        M3 dR = (delta_R * exp_so3(J_dR_bg * d_bg)).T() * T_fi_w.R * T_fj_w.R.T();                   // :157-158
        V3 r_dr = log_so3(dR);
        V3 r_dv = T_fi_w.R * (v_j - v_i - g * dtij) - (delta_v + J_dv_bg * d_bg + J_dv_ba * d_ba);   // :160-161
        V3 pj = T_fj_w.inverse().t, pi = T_fi_w.inverse().t;
        V3 r_dp = T_fi_w.R * (pj - pi - v_i * dtij - (0.5 * dtij * dtij) * g) -
                  (delta_p + J_dp_bg * d_bg + J_dp_ba * d_ba);                                        // :162-164
        Mat<9, 1> err;
        for (int k = 0; k < 3; k++) {
            err[k] = r_dr[k];
            err[3 + k] = r_dv[k];
            err[6 + k] = r_dp[k];
        }
        err = inf_sqrt * err;                                                                         // :169
        err.to(residuals);
        if (jacobians != nullptr) {
            M3 Jr_inv = inverse3(so3_rightJacobian(r_dr));
            if (jacobians[0] != nullptr) {                                                            // :174-187
                Mat<9, 6> J = Mat<9, 6>::Zero();
                V3 w_dfi = vec3(parameters[0][0], parameters[0][1], parameters[0][2]);
                M3 J_r_wdfi = so3_rightJacobian(w_dfi);
                J.setBlock(0, 0, Jr_inv * T_fj_w.R * J_r_wdfi);
                J.setBlock(3, 0, -(T_fi_w.R * skewMatrix(v_j - v_i - g * dtij) * J_r_wdfi));
                J.setBlock(6, 0, -(T_fi_w.R * skewMatrix(pj - v_i * dtij - (0.5 * dtij * dtij) * g) * J_r_wdfi));
                J.setBlock(6, 3, T_fi_w_base.R);
                J = inf_sqrt * J;
                J.to(jacobians[0]);
            }
            if (jacobians[1] != nullptr) {                                                            // :190-200
                Mat<9, 6> J = Mat<9, 6>::Zero();
                V3 w_dfj = vec3(parameters[1][0], parameters[1][1], parameters[1][2]);
                M3 J_r_wdfj = so3_rightJacobian(w_dfj);
                J.setBlock(0, 0, -(Jr_inv * T_fj_w.R * J_r_wdfj));
                J.setBlock(6, 0, -(T_fi_w.R * T_fj_w.R.T() * skewMatrix(T_fj_w.t) * T_fj_w.R * J_r_wdfj));
                J.setBlock(6, 3, -(T_fi_w.R * exp_so3(w_dfj).T()));
                J = inf_sqrt * J;
                J.to(jacobians[1]);
            }
            if (jacobians[2] != nullptr) {                                                            // :203-209
                Mat<9, 3> J = Mat<9, 3>::Zero();
                J.setBlock(3, 0, -T_fi_w.R);
                J.setBlock(6, 0, -(T_fi_w.R * dtij));
                J = inf_sqrt * J;
                J.to(jacobians[2]);
            }
            if (jacobians[3] != nullptr) {                                                            // :212-217
                Mat<9, 3> J = Mat<9, 3>::Zero();
                J.setBlock(3, 0, T_fi_w.R);
                J = inf_sqrt * J;
                J.to(jacobians[3]);
            }
            if (jacobians[4] != nullptr) {                                                            // :220-226
                Mat<9, 3> J = Mat<9, 3>::Zero();
                J.setBlock(3, 0, -J_dv_ba);
                J.setBlock(6, 0, -J_dp_ba);
                J = inf_sqrt * J;
                J.to(jacobians[4]);
            }
            if (jacobians[5] != nullptr) {                                                            // :229-237
                Mat<9, 3> J = Mat<9, 3>::Zero();
                J.setBlock(0, 0, -(Jr_inv * dR.T() * so3_rightJacobian(J_dR_bg * d_bg) * J_dR_bg));
                J.setBlock(3, 0, -J_dv_bg);
                J.setBlock(6, 0, -J_dp_bg);
                J = inf_sqrt * J;
                J.to(jacobians[5]);
            }
        }
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// IMUFactorInit — residuals.hpp:302-410, SizedCostFunction<9, 2,3,3,3,3,1>, the functor of AOptimizer::VIInit
// parameters: (w_x, w_y) of the gravity alignment R_w_i = Exp(w_x, w_y, 0), dv_i, dv_j, dba, dbg, log-scale lambda
// ---------------------------------------------------------------------------------------------
struct ImuFactorInit {
    Aff T_fi_w, T_fj_w;     // poses are NOT parameters here (:316-317)
    V3 v_i_base, v_j_base;
    double dtij;
    Mat<9, 9> cov;
    M3 delta_R;
    V3 delta_v, delta_p;
    M3 J_dR_bg, J_dv_ba, J_dv_bg, J_dp_ba, J_dp_bg;
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const {
        V3 w_w_i = vec3(parameters[0][0], parameters[0][1], 0);                                       // :309
        M3 R_w_i = exp_so3(w_w_i);
        V3 v_i = v_i_base + V3::From(parameters[1]);                                                  // :311-312
        V3 v_j = v_j_base + V3::From(parameters[2]);
        V3 d_ba = V3::From(parameters[3]);
        V3 d_bg = V3::From(parameters[4]);
        double lambda = parameters[5][0];
        V3 g = V3::From(kGravity);
        Mat<9, 9> inf_sqrt;
        if (!ImuFactor::InfSqrt(cov, inf_sqrt)) return false;                                         // :321-323
        M3 dR = (delta_R * exp_so3(J_dR_bg * d_bg)).T() * T_fi_w.R * T_fj_w.R.T();                    // :326-327
        V3 r_dr = log_so3(dR);
        M3 RiRw = T_fi_w.R * R_w_i;
        V3 a_v = (v_j - v_i) - g * dtij;
        V3 dpos = T_fj_w.inverse().t - T_fi_w.inverse().t;
        V3 a_p = std::exp(lambda) * dpos - v_i * dtij - (0.5 * dtij * dtij) * g;
        V3 r_dv = RiRw * a_v - (delta_v + J_dv_bg * d_bg + J_dv_ba * d_ba);                           // :329-330
        V3 r_dp = RiRw * a_p - (delta_p + J_dp_bg * d_bg + J_dp_ba * d_ba);                           // :331-335
        Mat<9, 1> err;
        for (int k = 0; k < 3; k++) {
            err[k] = r_dr[k];
            err[3 + k] = r_dv[k];
            err[6 + k] = r_dp[k];
        }
        err = inf_sqrt * err;                                                                         // :340
        err.to(residuals);
        if (jacobians != nullptr) {
            if (jacobians[0] != nullptr) {                                                            // :345-356
                Mat<9, 2> J = Mat<9, 2>::Zero();
                M3 Jr = so3_rightJacobian(w_w_i);
                M3 Bv = -(RiRw * skewMatrix(a_v) * Jr), Bp = -(RiRw * skewMatrix(a_p) * Jr);
                for (int r = 0; r < 3; r++)
                    for (int c = 0; c < 2; c++) {
                        J(3 + r, c) = Bv(r, c);
                        J(6 + r, c) = Bp(r, c);
                    }
                J = inf_sqrt * J;
                J.to(jacobians[0]);
            }
            if (jacobians[1] != nullptr) {                                                            // :359-365
                Mat<9, 3> J = Mat<9, 3>::Zero();
                J.setBlock(3, 0, -RiRw);
                J.setBlock(6, 0, -(RiRw * dtij));
                J = inf_sqrt * J;
                J.to(jacobians[1]);
            }
            if (jacobians[2] != nullptr) {                                                            // :368-373
                Mat<9, 3> J = Mat<9, 3>::Zero();
                J.setBlock(3, 0, RiRw);
                J = inf_sqrt * J;
                J.to(jacobians[2]);
            }
            if (jacobians[3] != nullptr) {                                                            // :376-382
                Mat<9, 3> J = Mat<9, 3>::Zero();
                J.setBlock(3, 0, -J_dv_ba);
                J.setBlock(6, 0, -J_dp_ba);
                J = inf_sqrt * J;
                J.to(jacobians[3]);
            }
            if (jacobians[4] != nullptr) {                                                            // :385-393
                Mat<9, 3> J = Mat<9, 3>::Zero();
                J.setBlock(0, 0, -(inverse3(so3_rightJacobian(r_dr)) * dR.T() * so3_rightJacobian(J_dR_bg * d_bg) * J_dR_bg));
                J.setBlock(3, 0, -J_dv_bg);
                J.setBlock(6, 0, -J_dp_bg);
                J = inf_sqrt * J;
                J.to(jacobians[4]);
            }
            if (jacobians[5] != nullptr) {                                                            // :396-402
                Mat<9, 1> J = Mat<9, 1>::Zero();
                V3 s = RiRw * dpos;   // NOTE: the reference's derivative omits the exp(lambda) factor (:399-400); reproduced
                for (int r = 0; r < 3; r++) J[6 + r] = s[r];
                J = inf_sqrt * J;
                J.to(jacobians[5]);
            }
        }
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// IMUBiasFactor — residuals.hpp:247-300, SizedCostFunction<6, 3,3,3,3>
// parameters: dba_i, dbg_i, dba_j, dbg_j
// ---------------------------------------------------------------------------------------------
struct ImuBiasFactor {
    V3 ba_i, bg_i, ba_j, bg_j;
    double dtij, sigma_ba, sigma_bg; // imu_i->getbAccNoise(), getbGyrNoise()
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const {
        V3 d_bai = V3::From(parameters[0]), d_bgi = V3::From(parameters[1]);
        V3 d_baj = V3::From(parameters[2]), d_bgj = V3::From(parameters[3]);
        double sigma2_dba = dtij * sigma_ba * sigma_ba;                 // :259
        double wa = 1 / std::sqrt(sigma2_dba);
        double sigma2_dbg = dtij * sigma_bg * sigma_bg;                 // :261
        double wg = 1 / std::sqrt(sigma2_dbg);
        V3 ea = wa * (ba_j + d_baj - ba_i - d_bai);                     // :265
        V3 eg = wg * (bg_j + d_bgj - bg_i - d_bgi);                     // :266
        for (int k = 0; k < 3; k++) {
            residuals[k] = ea[k];
            residuals[3 + k] = eg[k];
        }
        if (jacobians != nullptr) {                                     // :268-293
            const double sgn[4] = {-1, -1, 1, 1};
            for (int b = 0; b < 4; b++) {
                if (jacobians[b] == nullptr) continue;
                Mat<6, 3> J = Mat<6, 3>::Zero();
                bool is_ba = (b == 0 || b == 2);
                for (int k = 0; k < 3; k++) J((is_ba ? 0 : 3) + k, k) = sgn[b] * (is_ba ? wa : wg);
                J.to(jacobians[b]);
            }
        }
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// PosePriordx — residuals.hpp:601-632, SizedCostFunction<6, 6>
// ---------------------------------------------------------------------------------------------
inline void pose_prior_core(const Aff &T_base, const Aff &T_prior, const double *dx, Mat<6, 1> &err, Mat<6, 6> *J) {
    Aff T = T_base * se3_doubleVec6dtoRT(dx);                           // :609
    err = se3_RTtoVec6d(T * T_prior.inverse());                         // :610
    if (J) {
        *J = Mat<6, 6>::Identity();                                     // :614
        V3 dw = vec3(dx[0], dx[1], dx[2]);
        V3 w = log_so3(T.R * T_prior.R.T());                            // :617
        J->setBlock(0, 0, inverse3(so3_rightJacobian(w)) * T_prior.R * so3_rightJacobian(dw));        // :618-619
        J->setBlock(3, 0, T.R * skewMatrix(T_prior.R.T() * T_prior.t) * so3_rightJacobian(dw));       // :620-622
        J->setBlock(3, 3, T_base.R);                                    // :623
    }
}
struct PosePrior {
    Aff T, T_prior;
    Mat<6, 6> sqrt_inf;
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const {
        Mat<6, 1> e;
        Mat<6, 6> J;
        bool wantJ = jacobians != nullptr && jacobians[0] != nullptr;
        pose_prior_core(T, T_prior, parameters[0], e, wantJ ? &J : nullptr);
        (sqrt_inf * e).to(residuals);
        if (wantJ) (sqrt_inf * J).to(jacobians[0]);                     // :624
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// IMUPriordx — residuals.hpp:634-700, SizedCostFunction<15, 6,3,3,3>
// ---------------------------------------------------------------------------------------------
struct ImuPrior {
    Aff T, T_prior;
    V3 v, v_prior, ba, ba_prior, bg, bg_prior;
    Mat<15, 15> sqrt_inf;
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const {
        Mat<6, 1> e6;
        Mat<6, 6> J6;
        bool wantJ0 = jacobians != nullptr && jacobians[0] != nullptr;
        pose_prior_core(T, T_prior, parameters[0], e6, wantJ0 ? &J6 : nullptr);
        Mat<15, 1> err;
        for (int k = 0; k < 6; k++) err[k] = e6[k];
        for (int k = 0; k < 3; k++) {
            err[6 + k] = v[k] + parameters[1][k] - v_prior[k];          // :653
            err[9 + k] = ba[k] + parameters[2][k] - ba_prior[k];        // :654
            err[12 + k] = bg[k] + parameters[3][k] - bg_prior[k];       // :655
        }
        err = sqrt_inf * err;                                           // :656
        err.to(residuals);
        if (jacobians != nullptr) {
            if (jacobians[0] != nullptr) {                              // :660-674
                Mat<15, 6> J = Mat<15, 6>::Zero();
                // setZero() then blocks (0,0),(3,0),(3,3): the identity of PosePriordx is NOT kept here,
                // which is the same thing because (0,3) is zero in both.
                J.setBlock(0, 0, J6.block<3, 3>(0, 0));
                J.setBlock(3, 0, J6.block<3, 3>(3, 0));
                J.setBlock(3, 3, J6.block<3, 3>(3, 3));
                J = sqrt_inf * J;
                J.to(jacobians[0]);
            }
            // NOTE reference quirk (:676-692): the v/ba/bg Jacobians are NOT multiplied by sqrt_inf.
            for (int b = 1; b < 4; b++) {
                if (jacobians[b] == nullptr) continue;
                Mat<15, 3> J = Mat<15, 3>::Zero();
                J.setBlock(3 + 3 * b, 0, M3::Identity());
                J.to(jacobians[b]);
            }
        }
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// PoseToLandmarkFactor — residuals.hpp:561-599, SizedCostFunction<3, 6,3>
// ---------------------------------------------------------------------------------------------
struct PoseToLandmark {
    V3 delta, t_w_lmk;
    Aff T_f_w;
    M3 sqrt_inf;
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const {
        Aff T = T_f_w * se3_doubleVec6dtoRT(parameters[0]);             // :572
        V3 t = t_w_lmk + V3::From(parameters[1]);                       // :573
        (sqrt_inf * (T * t - delta)).to(residuals);                     // :575
        if (jacobians != nullptr) {
            if (jacobians[0] != nullptr) {                              // :579-586
                V3 w_df = vec3(parameters[0][0], parameters[0][1], parameters[0][2]);
                M3 dR = exp_so3(w_df);
                Mat<3, 6> J;
                J.setBlock(0, 0, sqrt_inf * T_f_w.R * (-(dR * skewMatrix(t) * so3_rightJacobian(w_df))));
                J.setBlock(0, 3, sqrt_inf * T_f_w.R);
                J.to(jacobians[0]);
            }
            if (jacobians[1] != nullptr) (sqrt_inf * T.R).to(jacobians[1]); // :588-591
        }
        return true;
    }
};

// Landmark3DPrior — residuals.hpp:506-526, SizedCostFunction<3,3>
struct LandmarkPrior {
    V3 prior, lmk;
    M3 sqrt_inf;
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const {
        (sqrt_inf * (lmk + V3::From(parameters[0]) - prior)).to(residuals); // :514
        if (jacobians != nullptr && jacobians[0] != nullptr) sqrt_inf.to(jacobians[0]);
        return true;
    }
};

// LandmarkToLandmarkFactor — residuals.hpp:528-559, SizedCostFunction<3,3,3>
struct LandmarkToLandmark {
    V3 delta, lmk0, lmk1;
    M3 sqrt_inf;
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const {
        (sqrt_inf * ((lmk0 + V3::From(parameters[0])) - (lmk1 + V3::From(parameters[1])) - delta)).to(residuals); // :539-540
        if (jacobians != nullptr) {
            if (jacobians[0] != nullptr) sqrt_inf.to(jacobians[0]);
            if (jacobians[1] != nullptr) (-sqrt_inf).to(jacobians[1]);
        }
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// IMU::processIMU — data/sensors/IMU.cpp:5-91 (one pre-integration step) and
// IMU::biasDeltaCorrection — IMU.cpp:104-108
// ---------------------------------------------------------------------------------------------
struct ImuState {
    // measurement + state carried by one isae::IMU object and its frame
    V3 acc, gyr, ba, bg, v;
    Aff T_f_w;
    bool frame_is_kf;
    M3 delta_R;
    V3 delta_v, delta_p;
    Mat<9, 9> Sigma;
    M3 J_dR_bg, J_dv_ba, J_dv_bg, J_dp_ba, J_dp_bg;
};

// cur.acc/gyr must be set by the caller; everything else of `cur` is produced here.
// `dt_in` = (ts_cur - ts_last) seconds; kf_ba/kf_bg = _last_kf->getIMU()->getBa()/getBg();
// eta = (gyr_noise^2, x3, acc_noise^2 x3) * rate_hz  (IMU.h:39-41).
inline void process_imu(const ImuState &last, const V3 &kf_ba, const V3 &kf_bg, double dt_in, const double eta[6],
                        double rate_hz, ImuState &cur) {
    V3 g = V3::From(kGravity);
    cur.ba = last.ba;                                                   // IMU.cpp:17-18
    cur.bg = last.bg;
    double dt = dt_in;                                                  // :21-25
    if (dt > 1) dt = 1 / rate_hz;
    double dt22 = 0.5 * dt * dt;
    V3 dv = (last.acc - last.ba) * dt;                                  // :28
    V3 dp = (last.acc - last.ba) * dt22;                                // :29
    M3 dR = exp_so3((last.gyr - last.bg) * dt);                         // :30
    M3 Jrk = so3_rightJacobian((last.gyr - kf_bg) * dt);                // :31
    Aff T_w_fp = last.T_f_w.inverse();
    M3 R_w_fp = T_w_fp.R;                                               // :34
    cur.v = last.v + R_w_fp * dv + g * dt;                              // :35
    Aff T_w_f = T_w_fp;                                                 // :38-41
    T_w_f.R = R_w_fp * dR;
    T_w_f.t = T_w_f.t + last.v * dt + R_w_fp * dp + g * dt22;
    cur.T_f_w = T_w_f.inverse();
    Mat<9, 6> B = Mat<9, 6>::Zero();                                    // :44-47
    B.setBlock(0, 0, Jrk * dt);
    B.setBlock(3, 3, last.delta_R * dt);
    B.setBlock(6, 3, last.delta_R * dt22);
    Mat<6, 6> Eta = Mat<6, 6>::Zero();
    for (int i = 0; i < 6; i++) Eta(i, i) = eta[i];
    if (last.frame_is_kf) {                                             // :50-61
        cur.delta_p = dp;
        cur.delta_v = dv;
        cur.delta_R = dR;
        cur.Sigma = B * Eta * B.T();
        for (int i = 0; i < 3; i++) cur.Sigma(6 + i, 6 + i) += 0.0001 * dt;
        cur.J_dR_bg = -(Jrk * dt);
        cur.J_dv_ba = -(M3::Identity() * dt);
        cur.J_dv_bg = M3::Zero();
        cur.J_dp_ba = -(dt22 * M3::Identity());
        cur.J_dp_bg = M3::Zero();
    } else {                                                            // :63-88
        cur.delta_R = last.delta_R * dR;
        cur.delta_v = last.delta_v + last.delta_R * dv;
        cur.delta_p = last.delta_p + last.delta_v * dt + last.delta_R * dp;
        Mat<9, 9> A = Mat<9, 9>::Identity();
        M3 dR_dA = last.delta_R * skewMatrix(last.acc - kf_ba);
        A.setBlock(0, 0, dR.T());
        A.setBlock(3, 0, -(dR_dA * dt));
        A.setBlock(6, 0, -(dR_dA * dt22));
        A.setBlock(6, 3, M3::Identity() * dt);
        cur.Sigma = A * last.Sigma * A.T() + B * Eta * B.T();
        for (int i = 0; i < 3; i++) cur.Sigma(6 + i, 6 + i) += 0.0001 * dt;
        cur.J_dR_bg = dR.T() * last.J_dR_bg - Jrk * dt;
        cur.J_dv_ba = last.J_dv_ba - last.delta_R * dt;
        cur.J_dv_bg = last.J_dv_bg - dR_dA * last.J_dR_bg * dt;
        cur.J_dp_ba = last.J_dp_ba + last.J_dv_ba * dt - dt22 * last.delta_R;
        cur.J_dp_bg = last.J_dp_bg + last.J_dv_bg * dt - dt22 * (dR_dA * last.J_dR_bg);
    }
}

inline void bias_delta_correction(ImuState &s, const V3 &d_ba, const V3 &d_bg) { // IMU.cpp:104-108
    s.delta_p = s.delta_p + s.J_dp_ba * d_ba + s.J_dp_bg * d_bg;
    s.delta_v = s.delta_v + s.J_dv_ba * d_ba + s.J_dv_bg * d_bg;
    s.delta_R = s.delta_R * exp_so3(s.J_dR_bg * d_bg);
}

} // namespace orc

// TEST INFRASTRUCTURE — CPU oracle for the sliding-window BA/VIO solve. See oracle/README.md.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
//
// Restates (behaviourally, not textually):
//   * the problem build of AOptimizer::localMapVIOptimization / localMapBA
//     (cpp/src/optimizers/AOptimizer.cpp:299-446, AOptimizer.cpp:22-96,
//      AngularAdjustmentCERESAnalytic.cpp:212-339,341-486), driven by the flattened sdv_window;
//   * ceres::Solve with the options of AOptimizer.cpp:376-388. Ceres 2.2.0 (docker/Dockerfile:50) is a
//     third-party dependency ABSENT from /root/reference; its Levenberg-Marquardt trust-region loop
//     (internal/ceres/trust_region_minimizer.cc, levenberg_marquardt_strategy.cc,
//     trust_region_step_evaluator.cc, and the reduced-program preprocessing) is restated here from its
//     published algorithm.  PARITY PIN: the cost functors and IMU pre-integration are pinned against the
//     reference's own known-answer tests (tests/test_oracle_kats.py); the LM loop itself is pinned only
//     through the reference's end-to-end tolerances (imu_test.cpp:485-487, :566-567) => for windows with
//     visual factors parity with a real Ceres build is UNPINNED (no Ceres/Eigen in this image).
#include "../include/sdv.h"
#include "factors.hpp"
#include "marg.hpp"

#include <array>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <limits>
#include <vector>

using namespace orc;

namespace {

// ------------------------------------------------------------------------------------------------
// Problem description (what ceres::Problem holds after the reference's add*Residuals calls)
// ------------------------------------------------------------------------------------------------
enum Kind { K_VISUAL = 0, K_POSE_PRIOR, K_IMU, K_IMU_BIAS, K_MARG, K_IMU_PRIOR, K_P2L, K_LMK_PRIOR, K_L2L, K_IMU_INIT };

struct PBlock {
    int size = 0;
    int off = 0;         // offset into the state vector x
    bool constant = false;
    bool active = false; // present in Ceres' reduced program
    bool elim = false;   // eliminable landmark (Schur mode); dense otherwise
    int col = -1;        // column offset in the reduced program (oracle ordering: dense first, then eliminable)
};

struct RBlock {
    int kind, idx;
    int nres;
    int npb;
    std::vector<int> pb; // parameter block ids
    bool active = true;  // false => every parameter block constant => goes to fixed_cost
    size_t joff = 0;     // offset of its jacobian storage
    int roff = 0;        // offset of its residuals
};

struct Problem {
    const sdv_window *w;
    int F, L;
    std::vector<PBlock> pbs;
    std::vector<RBlock> rbs;
    int nx = 0;          // total state size (all blocks)
    int ncols = 0;       // reduced-program columns
    int ndense = 0;      // columns of non-eliminable blocks
    int nres = 0;        // residual rows in the reduced program
    size_t jsize = 0;
    std::vector<Mat<9, 9>> imu_inf_sqrt; // cached per IMU factor (constant across evaluations)
    // ids
    int pose_id(int f) const { return f; }
    int vel_id(int f) const { return F + f; }
    int ba_id(int f) const { return 2 * F + f; }
    int bg_id(int f) const { return 3 * F + f; }
    int lmk_id(int l) const { return 4 * F + l; }
    // AOptimizer::VIInit (AOptimizer.cpp:448-581): 0 = a window solve, 1 = VIInit with the scale constant, 2 = scale free,
    // 3 = every block of the functor free incl. the shared dba / dbg (the reference's own functor test, imu_test.cpp:498-541)
    int viinit = 0;
    int rwi_id() const { return 4 * F + L; }     // r_wi_par, size 2 (:467-468)
    int lam_id() const { return 4 * F + L + 1; } // lambda, size 1 (:480-483)
};

static double sigma_of_obs(const sdv_window *w, int o) {
    if (w->obs_sigma) return w->obs_sigma[o];
    if (w->factor_kind == SDV_FACTOR_ANGULAR) {
        const double *K = w->K + 4 * w->obs_cam[o];
        return 1.5 / ((K[0] + K[1]) / 2); // AngularAdjustmentCERESAnalytic.cpp:283, Camera.h:46
    }
    return 1.0; // BundleAdjustmentCERESAnalytic.h:47 default sigma
}

static void reduce_program(Problem &P) {
    // ---- Ceres reduced program: drop constant blocks, drop residual blocks with only constant blocks,
    //      drop parameter blocks no remaining residual block uses.
    for (auto &r : P.rbs) {
        r.active = false;
        for (int id : r.pb)
            if (!P.pbs[id].constant) r.active = true;
        if (r.active)
            for (int id : r.pb)
                if (!P.pbs[id].constant) P.pbs[id].active = true;
    }
    int col = 0;
    for (auto &p : P.pbs)
        if (p.active && !p.elim) {
            p.col = col;
            col += p.size;
        }
    P.ndense = col;
    for (auto &p : P.pbs)
        if (p.active && p.elim) {
            p.col = col;
            col += p.size;
        }
    P.ncols = col;
    size_t joff = 0;
    int roff = 0;
    for (auto &r : P.rbs) {
        if (!r.active) continue;
        r.joff = joff;
        r.roff = roff;
        int width = 0;
        for (int id : r.pb) width += P.pbs[id].size;
        joff += (size_t)r.nres * width;
        roff += r.nres;
    }
    P.jsize = joff;
    P.nres = roff;
}

static bool build_problem_viinit(const sdv_window *w, Problem &P);

static bool build_problem(const sdv_window *w, Problem &P, bool schur, int viinit = 0) {
    P.w = w;
    P.viinit = viinit;
    if (viinit) return build_problem_viinit(w, P);
    const int F = P.F = w->n_frames, L = P.L = w->n_lmks;
    P.pbs.assign(4 * F + L, PBlock());
    int off = 0;
    for (int f = 0; f < F; f++) {
        // fixed frames: (int)i > (int)(size - fixed - 1)  (AngularAdjustmentCERESAnalytic.cpp:234, AOptimizer.cpp:47)
        bool fixed = f > (F - w->n_fixed - 1);
        PBlock &p = P.pbs[P.pose_id(f)];
        p.size = 6;
        p.constant = fixed;
    }
    for (int k = 1; k <= 3; k++)
        for (int f = 0; f < F; f++) {
            bool fixed = f > (F - w->n_fixed - 1);
            PBlock &p = P.pbs[k * F + f];
            p.size = 3;
            p.constant = fixed;
        }
    for (int l = 0; l < L; l++) {
        PBlock &p = P.pbs[P.lmk_id(l)];
        p.size = 3;
        p.elim = schur;
        p.constant = w->landmarks_constant != 0; // single-frame solves: SetParameterBlockConstant, AngularAdjustmentCERESAnalytic.cpp:41-43
    }
    for (auto &p : P.pbs) {
        p.off = off;
        off += p.size;
    }
    P.nx = off;

    auto add = [&](int kind, int idx, int nres, std::initializer_list<int> pb) {
        RBlock r;
        r.kind = kind;
        r.idx = idx;
        r.nres = nres;
        r.pb.assign(pb);
        r.npb = (int)r.pb.size();
        P.rbs.push_back(r);
    };
    // pose priors (AngularAdjustmentCERESAnalytic.cpp:239-243)
    for (int f = 0; f < F; f++)
        if (w->has_prior && w->has_prior[f]) add(K_POSE_PRIOR, f, 6, {P.pose_id(f)});
    // visual factors (…:266-288), reference walk order
    for (int o = 0; o < w->n_obs; o++) add(K_VISUAL, o, 2, {P.pose_id(w->obs_frame[o]), P.lmk_id(w->obs_lmk[o])});
    // IMU + bias factors (AOptimizer.cpp:72-93)
    if (w->vio) {
        P.imu_inf_sqrt.resize(w->n_imu);
        for (int p = 0; p < w->n_imu; p++) {
            int i = w->imu_i[p], j = w->imu_j[p];
            add(K_IMU, p, 9, {P.pose_id(i), P.pose_id(j), P.vel_id(i), P.vel_id(j), P.ba_id(i), P.bg_id(i)});
            add(K_IMU_BIAS, p, 6, {P.ba_id(i), P.bg_id(i), P.ba_id(j), P.bg_id(j)});
            if (!ImuFactor::InfSqrt(Mat<9, 9>::From(w->imu_cov + 81 * p), P.imu_inf_sqrt[p])) return false;
        }
    }
    // marginalisation priors (AngularAdjustmentCERESAnalytic.cpp:341-486)
    if (w->dense_prior) {
        const sdv_dense_prior *dp = w->dense_prior;
        RBlock r;
        r.kind = K_MARG;
        r.idx = 0;
        r.nres = dp->n_full;
        if (dp->frame >= 0) {
            r.pb.push_back(P.pose_id(dp->frame));
            r.pb.push_back(P.vel_id(dp->frame));
            r.pb.push_back(P.ba_id(dp->frame));
            r.pb.push_back(P.bg_id(dp->frame));
        }
        for (int k = 0; k < dp->n_keep; k++) {
            if (dp->keep_col[k] < 0) continue; // marginalization.hpp:139
            r.pb.push_back(P.lmk_id(dp->keep_lmk[k]));
            P.pbs[P.lmk_id(dp->keep_lmk[k])].elim = false; // group 2: stays in the reduced system
        }
        r.npb = (int)r.pb.size();
        P.rbs.push_back(r);
    }
    if (w->sparse_prior) {
        const sdv_sparse_prior *sp = w->sparse_prior;
        if (sp->has_imu_prior) {
            int f = sp->frame;
            add(K_IMU_PRIOR, 0, 15, {P.pose_id(f), P.vel_id(f), P.ba_id(f), P.bg_id(f)});
            for (int k = 0; k < sp->n_p2l; k++) add(K_P2L, k, 3, {P.pose_id(f), P.lmk_id(sp->p2l_lmk[k])});
        }
        if (sp->has_lmk_prior) {
            add(K_LMK_PRIOR, 0, 3, {P.lmk_id(sp->lmk0)});
            P.pbs[P.lmk_id(sp->lmk0)].elim = false;
        }
        for (int k = 0; k < sp->n_l2l; k++) {
            add(K_L2L, k, 3, {P.lmk_id(sp->l2l_a[k]), P.lmk_id(sp->l2l_b[k])});
            P.pbs[P.lmk_id(sp->l2l_a[k])].elim = false;
            P.pbs[P.lmk_id(sp->l2l_b[k])].elim = false;
        }
    }
    reduce_program(P);
    return true;
}

// The problem AOptimizer::VIInit builds (AOptimizer.cpp:448-529): one velocity block per frame with an IMU (:459-464), the 2-dof
// gravity alignment (:467-468), the log-scale (:480-483, constant unless optim_scale), ONE shared dba / dbg pair that is set
// constant (:472-477) and therefore stays at zero; one IMUFactorInit per (getLastKF(), frame) pair (:485-502, no dt test).  The
// two Landmark3DPrior blocks on dba / dbg (:504-515) only touch constant blocks (Ceres drops them; their residual is zero).
static bool build_problem_viinit(const sdv_window *w, Problem &P) {
    const int F = P.F = w->n_frames, L = P.L = w->n_lmks;
    P.pbs.assign(4 * F + L + 2, PBlock());
    for (int f = 0; f < F; f++) {
        P.pbs[P.pose_id(f)].size = 6;
        P.pbs[P.pose_id(f)].constant = true; // poses are not parameters of VIInit
        for (int k = 1; k <= 3; k++) {
            P.pbs[k * F + f].size = 3;
            P.pbs[k * F + f].constant = k != 1 && !(P.viinit == 3 && f == 0); // (mode 3: blocks ba_id(0) / bg_id(0) stand for the shared dba / dbg)
        }
    }
    for (int l = 0; l < L; l++) {
        P.pbs[P.lmk_id(l)].size = 3;
        P.pbs[P.lmk_id(l)].constant = true;
    }
    P.pbs[P.rwi_id()].size = 2;
    P.pbs[P.lam_id()].size = 1;
    P.pbs[P.lam_id()].constant = P.viinit == 1;
    int off = 0;
    for (auto &p : P.pbs) {
        p.off = off;
        off += p.size;
    }
    P.nx = off;
    P.imu_inf_sqrt.resize(w->n_imu);
    for (int p = 0; p < w->n_imu; p++) {
        RBlock r;
        r.kind = K_IMU_INIT;
        r.idx = p;
        r.nres = 9;
        if (P.viinit == 3) r.pb = {P.rwi_id(), P.vel_id(w->imu_i[p]), P.vel_id(w->imu_j[p]), P.ba_id(0), P.bg_id(0), P.lam_id()};
        else r.pb = {P.rwi_id(), P.vel_id(w->imu_i[p]), P.vel_id(w->imu_j[p]), P.lam_id()};
        r.npb = (int)r.pb.size();
        P.rbs.push_back(r);
        if (!ImuFactor::InfSqrt(Mat<9, 9>::From(w->imu_cov + 81 * p), P.imu_inf_sqrt[p])) return false;
    }
    reduce_program(P);
    return true;
}

// Evaluate one residual block at state x. jac (may be null) = concatenated row-major blocks, one per parameter.
static bool eval_block(const Problem &P, const RBlock &rb, const double *x, double *res, double *jac) {
    const sdv_window *w = P.w;
    const double *par[64];
    double *jp[64];
    std::vector<const double *> par_big;
    std::vector<double *> jp_big;
    const double **pp = par;
    double **jj = jp;
    if (rb.npb > 64) {
        par_big.resize(rb.npb);
        jp_big.resize(rb.npb);
        pp = par_big.data();
        jj = jp_big.data();
    }
    size_t o = 0;
    for (int k = 0; k < rb.npb; k++) {
        const PBlock &pb = P.pbs[rb.pb[k]];
        pp[k] = x + pb.off;
        jj[k] = jac ? jac + o : nullptr;
        o += (size_t)rb.nres * pb.size;
    }
    double **J = jac ? jj : nullptr;
    switch (rb.kind) {
    case K_VISUAL: {
        int ob = rb.idx, f = w->obs_frame[ob], c = w->obs_cam[ob], l = w->obs_lmk[ob];
        if (w->factor_kind == SDV_FACTOR_ANGULAR) {
            AngularErr e;
            e.bearing = V3::From(w->obs_bearing + 3 * ob);
            e.T_s_f = Aff::From(w->T_s_f + 12 * c);
            e.T_f_w = Aff::From(w->T_f_w + 12 * f);
            e.t_w_lmk = V3::From(w->lmk_t + 3 * l);
            e.sigma = sigma_of_obs(w, ob);
            return e.Evaluate(pp, res, J);
        } else {
            ReprojErr e;
            e.p2d[0] = w->obs_uv[2 * ob];
            e.p2d[1] = w->obs_uv[2 * ob + 1];
            for (int k = 0; k < 4; k++) e.K[k] = w->K[4 * c + k];
            e.T_s_f = Aff::From(w->T_s_f + 12 * c);
            e.T_f_w_base = Aff::From(w->T_f_w + 12 * f);
            e.t_w_lmk = V3::From(w->lmk_t + 3 * l);
            e.sigma = sigma_of_obs(w, ob);
            return e.Evaluate(pp, res, J);
        }
    }
    case K_POSE_PRIOR: {
        int f = rb.idx;
        PosePrior e;
        e.T = Aff::From(w->T_f_w + 12 * f);
        e.T_prior = Aff::From(w->T_prior + 12 * f);
        e.sqrt_inf = Mat<6, 6>::Zero();
        for (int k = 0; k < 6; k++) e.sqrt_inf(k, k) = w->inf_prior[6 * f + k];
        return e.Evaluate(pp, res, J);
    }
    case K_IMU: {
        int p = rb.idx, i = w->imu_i[p], j = w->imu_j[p];
        ImuFactor e;
        e.T_fi_w_base = Aff::From(w->T_f_w + 12 * i);
        e.T_fj_w_base = Aff::From(w->T_f_w + 12 * j);
        e.v_i_base = V3::From(w->v + 3 * i);
        e.v_j_base = V3::From(w->v + 3 * j);
        e.dtij = w->imu_dt[p];
        e.cov = Mat<9, 9>::From(w->imu_cov + 81 * p);
        e.delta_R = M3::From(w->imu_dR + 9 * p);
        e.delta_v = V3::From(w->imu_dv + 3 * p);
        e.delta_p = V3::From(w->imu_dp + 3 * p);
        e.J_dR_bg = M3::From(w->imu_J_dR_bg + 9 * p);
        e.J_dv_ba = M3::From(w->imu_J_dv_ba + 9 * p);
        e.J_dv_bg = M3::From(w->imu_J_dv_bg + 9 * p);
        e.J_dp_ba = M3::From(w->imu_J_dp_ba + 9 * p);
        e.J_dp_bg = M3::From(w->imu_J_dp_bg + 9 * p);
        return e.Evaluate(pp, res, J);
    }
    case K_IMU_INIT: {
        int p = rb.idx, i = w->imu_i[p], j = w->imu_j[p];
        ImuFactorInit e;
        e.T_fi_w = Aff::From(w->T_f_w + 12 * i);
        e.T_fj_w = Aff::From(w->T_f_w + 12 * j);
        e.v_i_base = V3::From(w->v + 3 * i);
        e.v_j_base = V3::From(w->v + 3 * j);
        e.dtij = w->imu_dt[p];
        e.cov = Mat<9, 9>::From(w->imu_cov + 81 * p);
        e.delta_R = M3::From(w->imu_dR + 9 * p);
        e.delta_v = V3::From(w->imu_dv + 3 * p);
        e.delta_p = V3::From(w->imu_dp + 3 * p);
        e.J_dR_bg = M3::From(w->imu_J_dR_bg + 9 * p);
        e.J_dv_ba = M3::From(w->imu_J_dv_ba + 9 * p);
        e.J_dv_bg = M3::From(w->imu_J_dv_bg + 9 * p);
        e.J_dp_ba = M3::From(w->imu_J_dp_ba + 9 * p);
        e.J_dp_bg = M3::From(w->imu_J_dp_bg + 9 * p);
        if (rb.npb == 6) return e.Evaluate(pp, res, J); // every block a parameter (imu_test.cpp:498-541)
        // the shared dba / dbg blocks are constant at zero: Ceres passes their values and a NULL Jacobian pointer
        static const double zero3[3] = {0, 0, 0};
        const double *p6[6] = {pp[0], pp[1], pp[2], zero3, zero3, pp[3]};
        double *j6[6] = {J ? jj[0] : nullptr, J ? jj[1] : nullptr, J ? jj[2] : nullptr, nullptr, nullptr, J ? jj[3] : nullptr};
        return e.Evaluate(p6, res, J ? j6 : nullptr);
    }
    case K_IMU_BIAS: {
        int p = rb.idx, i = w->imu_i[p], j = w->imu_j[p];
        ImuBiasFactor e;
        e.ba_i = V3::From(w->ba + 3 * i);
        e.bg_i = V3::From(w->bg + 3 * i);
        e.ba_j = V3::From(w->ba + 3 * j);
        e.bg_j = V3::From(w->bg + 3 * j);
        e.dtij = w->imu_dt[p];
        e.sigma_ba = w->imu_sigma_ba[p];
        e.sigma_bg = w->imu_sigma_bg[p];
        return e.Evaluate(pp, res, J);
    }
    case K_MARG: { // MarginalizationFactor::Evaluate, marginalization.hpp:113-215
        const sdv_dense_prior *dp = w->dense_prior;
        std::vector<double> dx(dp->n, 0.0);
        std::vector<int> cols;
        int b = 0;
        if (dp->frame >= 0) {
            const int offs[4] = {0, 6, 9, 12};
            const int sz[4] = {6, 3, 3, 3};
            for (int k = 0; k < 4; k++, b++) {
                for (int q = 0; q < sz[k]; q++) dx[dp->frame_col + offs[k] + q] = pp[b][q];
                cols.push_back(dp->frame_col + offs[k]);
            }
        }
        for (int k = 0; k < dp->n_keep; k++) {
            if (dp->keep_col[k] < 0) continue;
            for (int q = 0; q < 3; q++) dx[dp->keep_col[k] + q] = pp[b][q];
            cols.push_back(dp->keep_col[k]);
            b++;
        }
        for (int i = 0; i < dp->n_full; i++) {
            double s = dp->r0[i];
            const double *row = dp->J + (size_t)i * dp->n;
            for (int q = 0; q < dp->n; q++) s += row[q] * dx[q];
            res[i] = s;
        }
        if (J) {
            for (int k = 0; k < rb.npb; k++) {
                int sz = P.pbs[rb.pb[k]].size;
                for (int i = 0; i < dp->n_full; i++)
                    for (int q = 0; q < sz; q++) jj[k][i * sz + q] = dp->J[(size_t)i * dp->n + cols[k] + q];
            }
        }
        return true;
    }
    case K_IMU_PRIOR: {
        const sdv_sparse_prior *sp = w->sparse_prior;
        int f = sp->frame;
        ImuPrior e;
        e.T = Aff::From(w->T_f_w + 12 * f);
        e.T_prior = Aff::From(sp->T_prior);
        e.v = V3::From(w->v + 3 * f);
        e.ba = V3::From(w->ba + 3 * f);
        e.bg = V3::From(w->bg + 3 * f);
        e.v_prior = V3::From(sp->v_prior);
        e.ba_prior = V3::From(sp->ba_prior);
        e.bg_prior = V3::From(sp->bg_prior);
        e.sqrt_inf = Mat<15, 15>::From(sp->imu_sqrt_inf);
        return e.Evaluate(pp, res, J);
    }
    case K_P2L: {
        const sdv_sparse_prior *sp = w->sparse_prior;
        int k = rb.idx;
        PoseToLandmark e;
        e.delta = V3::From(sp->p2l_delta + 3 * k);
        e.t_w_lmk = V3::From(w->lmk_t + 3 * sp->p2l_lmk[k]);
        e.T_f_w = Aff::From(w->T_f_w + 12 * sp->frame);
        e.sqrt_inf = M3::From(sp->p2l_sqrt_inf + 9 * k);
        return e.Evaluate(pp, res, J);
    }
    case K_LMK_PRIOR: {
        const sdv_sparse_prior *sp = w->sparse_prior;
        LandmarkPrior e;
        e.prior = V3::From(sp->lmk_prior);
        e.lmk = V3::From(w->lmk_t + 3 * sp->lmk0);
        e.sqrt_inf = M3::From(sp->lmk_sqrt_inf);
        return e.Evaluate(pp, res, J);
    }
    case K_L2L: {
        const sdv_sparse_prior *sp = w->sparse_prior;
        int k = rb.idx;
        LandmarkToLandmark e;
        e.delta = V3::From(sp->l2l_delta + 3 * k);
        e.lmk0 = V3::From(w->lmk_t + 3 * sp->l2l_a[k]);
        e.lmk1 = V3::From(w->lmk_t + 3 * sp->l2l_b[k]);
        e.sqrt_inf = M3::From(sp->l2l_sqrt_inf + 9 * k);
        return e.Evaluate(pp, res, J);
    }
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// parallel-for helper (std::thread; the reference itself only uses ceres num_threads = 4)
// ------------------------------------------------------------------------------------------------
template <class Fn> static void parallel_for(int n, int nthreads, Fn fn) {
    if (nthreads <= 1 || n < 2 * nthreads) {
        fn(0, n, 0);
        return;
    }
    std::vector<std::thread> th;
    int chunk = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        int a = t * chunk, b = std::min(n, a + chunk);
        if (a >= b) break;
        th.emplace_back([=] { fn(a, b, t); });
    }
    for (auto &t : th) t.join();
}

// ceres::HuberLoss(a)::Evaluate followed by ceres::internal::Corrector (Ceres 2.2, not part of /root/reference: loss_function.cc,
// corrector.cc, residual_block.cc as remembered): rho(s) = s for s <= a^2, 2 a sqrt(s) - a^2 beyond; rho'' <= 0 always, so the
// corrector scales residuals AND Jacobian by sqrt(rho') and the block's cost is rho(s) / 2.  Returns rho(s).
static double huber_correct(double a, int nres, double *r, double *jac, int jcount) {
    double s = 0;
    for (int k = 0; k < nres; k++) s += r[k] * r[k];
    const double b = a * a;
    double rho0 = s, rho1 = 1.0;
    if (s > b) {
        const double rr = std::sqrt(s);
        rho0 = 2.0 * a * rr - b;
        rho1 = std::max(std::numeric_limits<double>::min(), a / rr);
    }
    const double sc = std::sqrt(rho1);
    for (int k = 0; k < nres; k++) r[k] *= sc;
    if (jac)
        for (int k = 0; k < jcount; k++) jac[k] *= sc;
    return rho0;
}

// Evaluate all active blocks. Returns cost = 1/2 sum rho(r^2) over active blocks; residuals/jacobians stored (robustified).
static bool evaluate(const Problem &P, const double *x, double *cost, double *residuals, double *jac, int nthreads) {
    int n = (int)P.rbs.size();
    std::vector<double> partial(std::max(1, nthreads), 0.0);
    std::atomic<bool> ok{true};
    parallel_for(n, nthreads, [&](int a, int b, int t) {
        double c = 0;
        std::vector<double> tmp;
        for (int i = a; i < b; i++) {
            const RBlock &rb = P.rbs[i];
            if (!rb.active) continue;
            double *r = residuals + rb.roff;
            if (!eval_block(P, rb, x, r, jac ? jac + rb.joff : nullptr)) ok = false;
            if (rb.kind == K_VISUAL && P.w->visual_loss_huber_a > 0) {
                int width = 0;
                for (int id : rb.pb) width += P.pbs[id].size;
                c += huber_correct(P.w->visual_loss_huber_a, rb.nres, r, jac ? jac + rb.joff : nullptr, rb.nres * width);
            } else
                for (int k = 0; k < rb.nres; k++) c += r[k] * r[k];
        }
        partial[t] = c;
    });
    double c = 0;
    for (double p : partial) c += p;
    *cost = 0.5 * c;
    return ok;
}

static double fixed_cost_of(const Problem &P, const double *x) {
    double c = 0;
    std::vector<double> r;
    for (auto &rb : P.rbs) {
        if (rb.active) continue;
        r.assign(rb.nres, 0.0);
        eval_block(P, rb, x, r.data(), nullptr);
        if (rb.kind == K_VISUAL && P.w->visual_loss_huber_a > 0) c += huber_correct(P.w->visual_loss_huber_a, rb.nres, r.data(), nullptr, 0);
        else
            for (double v : r) c += v * v;
    }
    return 0.5 * c;
}

// Dense in-place lower Cholesky (row-major n x n, only lower triangle referenced / written). Blocked, threaded
// trailing update. Returns false when a pivot is not positive.
static bool dense_cholesky(double *A, int n, int nthreads) {
    const int NB = 48;
    for (int k = 0; k < n; k += NB) {
        int kb = std::min(NB, n - k);
        // factor diagonal block
        for (int j = k; j < k + kb; j++) {
            double s = A[(size_t)j * n + j];
            for (int q = k; q < j; q++) s -= A[(size_t)j * n + q] * A[(size_t)j * n + q];
            if (!(s > 0.0) || !std::isfinite(s)) return false;
            double d = std::sqrt(s);
            A[(size_t)j * n + j] = d;
            for (int i = j + 1; i < k + kb; i++) {
                double t = A[(size_t)i * n + j];
                for (int q = k; q < j; q++) t -= A[(size_t)i * n + q] * A[(size_t)j * n + q];
                A[(size_t)i * n + j] = t / d;
            }
        }
        int rest = n - (k + kb);
        if (rest <= 0) break;
        // panel solve: rows below
        parallel_for(rest, nthreads, [&](int a, int b, int) {
            for (int i = k + kb + a; i < k + kb + b; i++) {
                double *Ai = A + (size_t)i * n;
                for (int j = k; j < k + kb; j++) {
                    double t = Ai[j];
                    const double *Aj = A + (size_t)j * n;
                    for (int q = k; q < j; q++) t -= Ai[q] * Aj[q];
                    Ai[j] = t / Aj[j];
                }
            }
        });
        // trailing update (lower triangle)
        parallel_for(rest, nthreads, [&](int a, int b, int) {
            for (int i = k + kb + a; i < k + kb + b; i++) {
                double *Ai = A + (size_t)i * n;
                for (int j = k + kb; j <= i; j++) {
                    const double *Aj = A + (size_t)j * n;
                    double s = 0;
                    for (int q = k; q < k + kb; q++) s += Ai[q] * Aj[q];
                    Ai[j] -= s;
                }
            }
        });
    }
    return true;
}
static void cholesky_solve(const double *Lm, int n, double *b) {
    for (int i = 0; i < n; i++) {
        double s = b[i];
        const double *Li = Lm + (size_t)i * n;
        for (int q = 0; q < i; q++) s -= Li[q] * b[q];
        b[i] = s / Li[i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = b[i] / Lm[(size_t)i * n + i];
        b[i] = s;
        for (int q = 0; q < i; q++) b[q] -= Lm[(size_t)i * n + q] * s;
    }
}

// ------------------------------------------------------------------------------------------------
// Linear solve of the LM sub-problem:  min || J y - r ||^2 + || D y ||^2   (J already column-scaled)
//   mode 0: Schur-eliminate landmark blocks, dense Cholesky of the reduced system
//   mode 1: dense Cholesky of the full normal equations (what SPARSE_NORMAL_CHOLESKY computes, densely)
// ------------------------------------------------------------------------------------------------
struct LinSys {
    std::vector<double> S;       // reduced (or full) system
    std::vector<double> rhs;
    // landmark elimination workspace
    std::vector<int> lmk_first;  // per eliminable pblock -> list head into lmk_rb
    std::vector<std::vector<int>> lmk_rbs;
};

static bool solve_linear(const Problem &P, const double *jac, const double *res, const double *D, double *y, int mode,
                         int nthreads, std::vector<std::vector<int>> &elim_rbs, std::vector<int> &elim_ids) {
    const int N = P.ncols;
    if (mode == 1 || P.ndense == N) {
        // full dense normal equations
        std::vector<double> H((size_t)N * N, 0.0), g(N, 0.0);
        for (auto &rb : P.rbs) {
            if (!rb.active) continue;
            size_t o1 = 0;
            for (int a = 0; a < rb.npb; a++) {
                const PBlock &pa = P.pbs[rb.pb[a]];
                const double *Ja = jac + rb.joff + o1;
                o1 += (size_t)rb.nres * pa.size;
                if (!pa.active) continue;
                for (int r = 0; r < rb.nres; r++)
                    for (int i = 0; i < pa.size; i++) g[pa.col + i] += Ja[r * pa.size + i] * res[rb.roff + r];
                size_t o2 = 0;
                for (int b = 0; b < rb.npb; b++) {
                    const PBlock &pb = P.pbs[rb.pb[b]];
                    const double *Jb = jac + rb.joff + o2;
                    o2 += (size_t)rb.nres * pb.size;
                    if (!pb.active || pb.col > pa.col) continue;
                    for (int i = 0; i < pa.size; i++)
                        for (int j = 0; j < pb.size; j++) {
                            double s = 0;
                            for (int r = 0; r < rb.nres; r++) s += Ja[r * pa.size + i] * Jb[r * pb.size + j];
                            H[(size_t)(pa.col + i) * N + pb.col + j] += s;
                        }
                }
            }
        }
        for (int i = 0; i < N; i++) H[(size_t)i * N + i] += D[i] * D[i];
        if (!dense_cholesky(H.data(), N, nthreads)) return false;
        cholesky_solve(H.data(), N, g.data());
        for (int i = 0; i < N; i++) y[i] = g[i];
        return true;
    }
    // ---- Schur mode
    const int n = P.ndense;
    int T = std::max(1, nthreads);
    std::vector<std::vector<double>> St(T), gt(T);
    for (int t = 0; t < T; t++) {
        St[t].assign((size_t)n * n, 0.0);
        gt[t].assign(n, 0.0);
    }
    // blocks without an eliminable parameter -> directly into S (thread 0)
    {
        std::vector<double> &S = St[0], &g = gt[0];
        for (auto &rb : P.rbs) {
            if (!rb.active) continue;
            bool has_elim = false;
            for (int id : rb.pb)
                if (P.pbs[id].active && P.pbs[id].elim) has_elim = true;
            if (has_elim) continue;
            size_t o1 = 0;
            for (int a = 0; a < rb.npb; a++) {
                const PBlock &pa = P.pbs[rb.pb[a]];
                const double *Ja = jac + rb.joff + o1;
                o1 += (size_t)rb.nres * pa.size;
                if (!pa.active) continue;
                for (int r = 0; r < rb.nres; r++)
                    for (int i = 0; i < pa.size; i++) g[pa.col + i] += Ja[r * pa.size + i] * res[rb.roff + r];
                size_t o2 = 0;
                for (int b = 0; b < rb.npb; b++) {
                    const PBlock &pb = P.pbs[rb.pb[b]];
                    const double *Jb = jac + rb.joff + o2;
                    o2 += (size_t)rb.nres * pb.size;
                    if (!pb.active || pb.col > pa.col) continue;
                    for (int i = 0; i < pa.size; i++)
                        for (int j = 0; j < pb.size; j++) {
                            double s = 0;
                            for (int r = 0; r < rb.nres; r++) s += Ja[r * pa.size + i] * Jb[r * pb.size + j];
                            S[(size_t)(pa.col + i) * n + pb.col + j] += s;
                        }
                }
            }
        }
    }
    const int NE = (int)elim_ids.size();
    std::vector<double> Vinv((size_t)NE * 9), gl((size_t)NE * 3);
    std::atomic<bool> ok{true};
    parallel_for(NE, T, [&](int a0, int b0, int t) {
        std::vector<double> &S = St[t], &g = gt[t];
        struct Ent { int col, size; double W[18]; }; // W = Jp^T Jl  (size x 3), size<=6
        std::vector<Ent> ents;
        for (int e = a0; e < b0; e++) {
            const PBlock &pl = P.pbs[elim_ids[e]];
            M3 V = M3::Zero();
            V3 g_l = V3::Zero();
            ents.clear();
            for (int rbi : elim_rbs[e]) {
                const RBlock &rb = P.rbs[rbi];
                // locate blocks
                size_t o = 0;
                const double *Jl = nullptr;
                std::vector<std::pair<const PBlock *, const double *>> others;
                for (int a = 0; a < rb.npb; a++) {
                    const PBlock &pa = P.pbs[rb.pb[a]];
                    const double *Ja = jac + rb.joff + o;
                    o += (size_t)rb.nres * pa.size;
                    if (rb.pb[a] == elim_ids[e]) Jl = Ja;
                    else if (pa.active) others.push_back({&pa, Ja});
                }
                const double *r = res + rb.roff;
                for (int i = 0; i < 3; i++) {
                    for (int j = 0; j < 3; j++) {
                        double s = 0;
                        for (int q = 0; q < rb.nres; q++) s += Jl[q * 3 + i] * Jl[q * 3 + j];
                        V(i, j) += s;
                    }
                    double s = 0;
                    for (int q = 0; q < rb.nres; q++) s += Jl[q * 3 + i] * r[q];
                    g_l[i] += s;
                }
                for (auto &oth : others) {
                    const PBlock &pa = *oth.first;
                    const double *Ja = oth.second;
                    // H_pp and g_p contributions of this residual block
                    for (int q = 0; q < rb.nres; q++)
                        for (int i = 0; i < pa.size; i++) g[pa.col + i] += Ja[q * pa.size + i] * r[q];
                    for (auto &oth2 : others) {
                        const PBlock &pb = *oth2.first;
                        if (pb.col > pa.col) continue;
                        const double *Jb = oth2.second;
                        for (int i = 0; i < pa.size; i++)
                            for (int j = 0; j < pb.size; j++) {
                                double s = 0;
                                for (int q = 0; q < rb.nres; q++) s += Ja[q * pa.size + i] * Jb[q * pb.size + j];
                                S[(size_t)(pa.col + i) * n + pb.col + j] += s;
                            }
                    }
                    // W accumulation per distinct dense block
                    Ent *en = nullptr;
                    for (auto &x : ents)
                        if (x.col == pa.col) en = &x;
                    if (!en) {
                        Ent ne;
                        ne.col = pa.col;
                        ne.size = pa.size;
                        for (int z = 0; z < 18; z++) ne.W[z] = 0;
                        ents.push_back(ne);
                        en = &ents.back();
                    }
                    for (int i = 0; i < pa.size; i++)
                        for (int j = 0; j < 3; j++) {
                            double s = 0;
                            for (int q = 0; q < rb.nres; q++) s += Ja[q * pa.size + i] * Jl[q * 3 + j];
                            en->W[i * 3 + j] += s;
                        }
                }
            }
            for (int i = 0; i < 3; i++) V(i, i) += D[pl.col + i] * D[pl.col + i];
            M3 Lc;
            if (!choleskyL<3>(V, Lc)) {
                ok = false;
                continue;
            }
            M3 Vi = inverse3(V);
            Vi.to(&Vinv[(size_t)e * 9]);
            g_l.to(&gl[(size_t)e * 3]);
            // S -= W_a Vi W_b^T ; g -= W_a Vi g_l
            for (auto &ea : ents) {
                double Y[18];
                for (int i = 0; i < ea.size; i++)
                    for (int j = 0; j < 3; j++) {
                        double s = 0;
                        for (int q = 0; q < 3; q++) s += ea.W[i * 3 + q] * Vi(q, j);
                        Y[i * 3 + j] = s;
                    }
                for (int i = 0; i < ea.size; i++) {
                    double s = 0;
                    for (int q = 0; q < 3; q++) s += Y[i * 3 + q] * g_l[q];
                    g[ea.col + i] -= s;
                }
                for (auto &eb : ents) {
                    if (eb.col > ea.col) continue;
                    for (int i = 0; i < ea.size; i++)
                        for (int j = 0; j < eb.size; j++) {
                            double s = 0;
                            for (int q = 0; q < 3; q++) s += Y[i * 3 + q] * eb.W[j * 3 + q];
                            S[(size_t)(ea.col + i) * n + eb.col + j] -= s;
                        }
                }
            }
        }
    });
    if (!ok) return false;
    std::vector<double> &S = St[0], &g = gt[0];
    for (int t = 1; t < T; t++) {
        for (size_t i = 0; i < (size_t)n * n; i++) S[i] += St[t][i];
        for (int i = 0; i < n; i++) g[i] += gt[t][i];
    }
    for (int i = 0; i < n; i++) S[(size_t)i * n + i] += D[i] * D[i];
    if (n > 0) {
        if (!dense_cholesky(S.data(), n, nthreads)) return false;
        cholesky_solve(S.data(), n, g.data());
    }
    for (int i = 0; i < n; i++) y[i] = g[i];
    // back-substitution: y_l = V^-1 (g_l - sum_a W_a^T y_a)
    parallel_for(NE, T, [&](int a0, int b0, int) {
        for (int e = a0; e < b0; e++) {
            const PBlock &pl = P.pbs[elim_ids[e]];
            double acc[3] = {gl[(size_t)e * 3], gl[(size_t)e * 3 + 1], gl[(size_t)e * 3 + 2]};
            for (int rbi : elim_rbs[e]) {
                const RBlock &rb = P.rbs[rbi];
                size_t o = 0;
                const double *Jl = nullptr;
                for (int a = 0; a < rb.npb; a++) {
                    if (rb.pb[a] == elim_ids[e]) Jl = jac + rb.joff + o;
                    o += (size_t)rb.nres * P.pbs[rb.pb[a]].size;
                }
                o = 0;
                for (int a = 0; a < rb.npb; a++) {
                    const PBlock &pa = P.pbs[rb.pb[a]];
                    const double *Ja = jac + rb.joff + o;
                    o += (size_t)rb.nres * pa.size;
                    if (rb.pb[a] == elim_ids[e] || !pa.active) continue;
                    for (int q = 0; q < rb.nres; q++) {
                        double u = 0;
                        for (int i = 0; i < pa.size; i++) u += Ja[q * pa.size + i] * y[pa.col + i];
                        for (int j = 0; j < 3; j++) acc[j] -= Jl[q * 3 + j] * u;
                    }
                }
            }
            const double *Vi = &Vinv[(size_t)e * 9];
            for (int i = 0; i < 3; i++) y[pl.col + i] = Vi[i * 3] * acc[0] + Vi[i * 3 + 1] * acc[1] + Vi[i * 3 + 2] * acc[2];
        }
    });
    return true;
}

// ------------------------------------------------------------------------------------------------
// Ceres 2.2 TrustRegionMinimizer + LevenbergMarquardtStrategy, restated.
// ------------------------------------------------------------------------------------------------
static int solve_lm(const sdv_window *w, const sdv_config *cfg, sdv_delta *out, sdv_stats *st, int mode, int nthreads,
                    double *S_out, double *g_out, int viinit = 0, double *extra_out = nullptr) {
    Problem P;
    if (!build_problem(w, P, mode == 0, viinit)) return SDV_ERR_NUMERICAL_FAILURE;
    const int N = P.ncols;
    std::vector<std::vector<int>> elim_rbs;
    std::vector<int> elim_ids, elim_index(P.pbs.size(), -1);
    for (size_t id = 0; id < P.pbs.size(); id++)
        if (P.pbs[id].active && P.pbs[id].elim) {
            elim_index[id] = (int)elim_ids.size();
            elim_ids.push_back((int)id);
        }
    elim_rbs.resize(elim_ids.size());
    int n_active_rb = 0;
    for (size_t i = 0; i < P.rbs.size(); i++) {
        if (!P.rbs[i].active) continue;
        n_active_rb++;
        for (int id : P.rbs[i].pb)
            if (elim_index[id] >= 0) elim_rbs[elim_index[id]].push_back((int)i);
    }
    std::vector<double> x(P.nx, 0.0), cand(P.nx, 0.0);
    std::vector<double> res(P.nres), jac(P.jsize), cres(P.nres);
    std::vector<double> grad(N), scale(N, 1.0), diag(N), D(N), step(N), delta(N), model(P.nres);
    double x_cost = 0, cand_cost = 0;

    std::memset(st, 0, sizeof(*st));
    st->n_reduced = P.ndense;
    st->n_residual_blocks = n_active_rb;
    st->fixed_cost = fixed_cost_of(P, x.data());

    auto col_sq_norms = [&](double *o) {
        for (int i = 0; i < N; i++) o[i] = 0;
        for (auto &rb : P.rbs) {
            if (!rb.active) continue;
            size_t off = 0;
            for (int a = 0; a < rb.npb; a++) {
                const PBlock &pa = P.pbs[rb.pb[a]];
                const double *Ja = jac.data() + rb.joff + off;
                off += (size_t)rb.nres * pa.size;
                if (!pa.active) continue;
                for (int r = 0; r < rb.nres; r++)
                    for (int i = 0; i < pa.size; i++) o[pa.col + i] += Ja[r * pa.size + i] * Ja[r * pa.size + i];
            }
        }
    };
    // EvaluateGradientAndJacobian (trust_region_minimizer.cc)
    int iteration = 0;
    auto eval_grad_jac = [&]() -> bool {
        if (!evaluate(P, x.data(), &x_cost, res.data(), jac.data(), nthreads)) return false;
        for (int i = 0; i < N; i++) grad[i] = 0;
        for (auto &rb : P.rbs) {
            if (!rb.active) continue;
            size_t off = 0;
            for (int a = 0; a < rb.npb; a++) {
                const PBlock &pa = P.pbs[rb.pb[a]];
                const double *Ja = jac.data() + rb.joff + off;
                off += (size_t)rb.nres * pa.size;
                if (!pa.active) continue;
                for (int r = 0; r < rb.nres; r++)
                    for (int i = 0; i < pa.size; i++) grad[pa.col + i] += Ja[r * pa.size + i] * res[rb.roff + r];
            }
        }
        if (cfg->jacobi_scaling) {
            if (iteration == 0) {
                col_sq_norms(scale.data());
                for (int i = 0; i < N; i++) scale[i] = 1.0 / (1.0 + std::sqrt(scale[i]));
            }
            for (auto &rb : P.rbs) {
                if (!rb.active) continue;
                size_t off = 0;
                for (int a = 0; a < rb.npb; a++) {
                    const PBlock &pa = P.pbs[rb.pb[a]];
                    double *Ja = jac.data() + rb.joff + off;
                    off += (size_t)rb.nres * pa.size;
                    if (!pa.active) continue;
                    for (int r = 0; r < rb.nres; r++)
                        for (int i = 0; i < pa.size; i++) Ja[r * pa.size + i] *= scale[pa.col + i];
                }
            }
        }
        return true;
    };
    auto grad_max = [&]() {
        double m = 0;
        for (int i = 0; i < N; i++) m = std::max(m, std::fabs(grad[i]));
        return m;
    };
    auto x_norm_of = [&](const std::vector<double> &v) {
        double s = 0;
        for (auto &p : P.pbs)
            if (p.active)
                for (int i = 0; i < p.size; i++) s += v[p.off + i] * v[p.off + i];
        return std::sqrt(s);
    };

    // LM strategy state
    double radius = cfg->initial_trust_region_radius, decrease_factor = 2.0;
    bool reuse_diagonal = false;
    int num_consecutive_invalid = 0;
    bool atleast_one_successful_step = false;
    int term = SDV_TERM_NO_CONVERGENCE;
    int n_ok = 0, n_bad = 0;

    // IterationZero
    if (!eval_grad_jac()) return SDV_ERR_NUMERICAL_FAILURE;
    double x_norm = x_norm_of(x);
    st->initial_cost = x_cost;
    st->trace_cost[0] = x_cost;
    st->trace_radius[0] = radius;
    st->trace_accepted[0] = 1;
    bool done = false;
    if (grad_max() <= cfg->gradient_tolerance) {
        term = SDV_TERM_GRADIENT_TOLERANCE;
        done = true;
    }
    if (S_out || g_out) {
        // debugging/testing aid: export the UNDAMPED reduced system at x0 in unscaled variables is not needed by
        // the solver; tests use orc_reduced_system below instead.
    }
    while (!done) {
        // FinalizeIterationAndCheckIfMinimizerCanContinue: max iterations / min radius
        if (iteration >= (w->max_num_iterations > 0 ? w->max_num_iterations : cfg->max_num_iterations)) {
            term = SDV_TERM_NO_CONVERGENCE;
            break;
        }
        if (radius <= cfg->min_trust_region_radius) { // MinTrustRegionRadiusReached
            term = SDV_TERM_MIN_RADIUS;
            break;
        }
        iteration++;
        // ---- ComputeTrustRegionStep -> LevenbergMarquardtStrategy::ComputeStep
        if (!reuse_diagonal) {
            col_sq_norms(diag.data());
            for (int i = 0; i < N; i++) diag[i] = std::min(std::max(diag[i], cfg->min_lm_diagonal), cfg->max_lm_diagonal);
        }
        for (int i = 0; i < N; i++) D[i] = std::sqrt(diag[i] / radius);
        bool lin_ok = solve_linear(P, jac.data(), res.data(), D.data(), step.data(), mode, nthreads, elim_rbs, elim_ids);
        if (lin_ok)
            for (int i = 0; i < N; i++)
                if (!std::isfinite(step[i])) lin_ok = false;
        reuse_diagonal = true;
        bool step_valid = false;
        double model_cost_change = 0;
        if (lin_ok) {
            for (int i = 0; i < N; i++) step[i] = -step[i];
            // model_residuals = J * step
            for (int i = 0; i < P.nres; i++) model[i] = 0;
            for (auto &rb : P.rbs) {
                if (!rb.active) continue;
                size_t off = 0;
                for (int a = 0; a < rb.npb; a++) {
                    const PBlock &pa = P.pbs[rb.pb[a]];
                    const double *Ja = jac.data() + rb.joff + off;
                    off += (size_t)rb.nres * pa.size;
                    if (!pa.active) continue;
                    for (int r = 0; r < rb.nres; r++) {
                        double s = 0;
                        for (int i = 0; i < pa.size; i++) s += Ja[r * pa.size + i] * step[pa.col + i];
                        model[rb.roff + r] += s;
                    }
                }
            }
            for (int i = 0; i < P.nres; i++) model_cost_change -= model[i] * (res[i] + model[i] / 2.0);
            step_valid = model_cost_change > 0.0;
        }
        int ti = std::min(iteration, SDV_MAX_TRACE - 1);
        st->trace_model_change[ti] = model_cost_change;
        if (!step_valid) {
            // HandleInvalidStep
            num_consecutive_invalid++;
            n_bad++;
            st->trace_cost[ti] = x_cost;
            st->trace_accepted[ti] = -1;
            if (num_consecutive_invalid >= cfg->max_consecutive_invalid_steps) {
                term = SDV_TERM_FAILURE;
                st->trace_radius[ti] = radius;
                break;
            }
            radius *= 0.5; // StepIsInvalid
            reuse_diagonal = false;
            st->trace_radius[ti] = radius;
            continue;
        }
        num_consecutive_invalid = 0;
        for (int i = 0; i < N; i++) delta[i] = step[i] * scale[i];
        // ComputeCandidatePointAndEvaluateCost
        cand = x;
        for (auto &p : P.pbs)
            if (p.active)
                for (int i = 0; i < p.size; i++) cand[p.off + i] = x[p.off + i] + delta[p.col + i];
        if (!evaluate(P, cand.data(), &cand_cost, cres.data(), nullptr, nthreads)) cand_cost = 1e300;
        // ParameterToleranceReached
        {
            double sn = 0;
            for (auto &p : P.pbs)
                if (p.active)
                    for (int i = 0; i < p.size; i++) sn += (x[p.off + i] - cand[p.off + i]) * (x[p.off + i] - cand[p.off + i]);
            sn = std::sqrt(sn);
            double tol = cfg->parameter_tolerance * (x_norm + cfg->parameter_tolerance);
            if (atleast_one_successful_step && sn <= tol) {
                term = SDV_TERM_PARAMETER_TOLERANCE;
                st->trace_cost[ti] = x_cost;
                st->trace_radius[ti] = radius;
                break;
            }
        }
        // FunctionToleranceReached (candidate is NOT applied)
        {
            double cost_change = x_cost - cand_cost;
            if (std::fabs(cost_change) <= cfg->function_tolerance * x_cost) {
                term = SDV_TERM_FUNCTION_TOLERANCE;
                st->trace_cost[ti] = x_cost;
                st->trace_radius[ti] = radius;
                break;
            }
        }
        // IsStepSuccessful (monotonic step evaluator)
        double relative_decrease = (x_cost - cand_cost) / model_cost_change;
        if (relative_decrease > cfg->min_relative_decrease) {
            // HandleSuccessfulStep
            x = cand;
            x_norm = x_norm_of(x);
            if (!eval_grad_jac()) return SDV_ERR_NUMERICAL_FAILURE;
            radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3)); // StepAccepted
            radius = std::min(cfg->max_trust_region_radius, radius);
            decrease_factor = 2.0;
            reuse_diagonal = false;
            atleast_one_successful_step = true;
            n_ok++;
            st->trace_cost[ti] = x_cost;
            st->trace_radius[ti] = radius;
            st->trace_accepted[ti] = 1;
            if (grad_max() <= cfg->gradient_tolerance) {
                term = SDV_TERM_GRADIENT_TOLERANCE;
                break;
            }
        } else {
            radius = radius / decrease_factor; // StepRejected
            decrease_factor *= 2.0;
            reuse_diagonal = true;
            n_bad++;
            st->trace_cost[ti] = x_cost;
            st->trace_radius[ti] = radius;
            st->trace_accepted[ti] = 0;
        }
    }
    st->iterations = iteration;
    st->termination = term;
    st->num_successful_steps = n_ok;
    st->num_unsuccessful_steps = n_bad;
    st->final_cost = x_cost;
    st->final_radius = radius;
    // write the parameter blocks out
    for (int f = 0; f < P.F; f++) {
        for (int k = 0; k < 6; k++) out->dpose[6 * f + k] = x[P.pbs[P.pose_id(f)].off + k];
        for (int k = 0; k < 3; k++) {
            if (out->dv) out->dv[3 * f + k] = x[P.pbs[P.vel_id(f)].off + k];
            if (out->dba) out->dba[3 * f + k] = x[P.pbs[P.ba_id(f)].off + k];
            if (out->dbg) out->dbg[3 * f + k] = x[P.pbs[P.bg_id(f)].off + k];
        }
    }
    for (int l = 0; l < P.L; l++)
        for (int k = 0; k < 3; k++) out->dlmk[3 * l + k] = x[P.pbs[P.lmk_id(l)].off + k];
    if (P.viinit && extra_out) {
        extra_out[0] = x[P.pbs[P.rwi_id()].off];
        extra_out[1] = x[P.pbs[P.rwi_id()].off + 1];
        extra_out[2] = x[P.pbs[P.lam_id()].off];
    }
    return term == SDV_TERM_FAILURE ? SDV_ERR_NUMERICAL_FAILURE : SDV_OK;
}

static void fill_x(const Problem &P, const sdv_delta *xd, std::vector<double> &x) {
    x.assign(P.nx, 0.0);
    if (!xd) return;
    for (int f = 0; f < P.F; f++) {
        if (xd->dpose)
            for (int k = 0; k < 6; k++) x[P.pbs[P.pose_id(f)].off + k] = xd->dpose[6 * f + k];
        for (int k = 0; k < 3; k++) {
            if (xd->dv) x[P.pbs[P.vel_id(f)].off + k] = xd->dv[3 * f + k];
            if (xd->dba) x[P.pbs[P.ba_id(f)].off + k] = xd->dba[3 * f + k];
            if (xd->dbg) x[P.pbs[P.bg_id(f)].off + k] = xd->dbg[3 * f + k];
        }
    }
    if (xd->dlmk)
        for (int l = 0; l < P.L; l++)
            for (int k = 0; k < 3; k++) x[P.pbs[P.lmk_id(l)].off + k] = xd->dlmk[3 * l + k];
}

} // namespace

// =================================================================================================
// C interface (loaded with ctypes from tests/ and bench.py's CPU-baseline leg)
// =================================================================================================
extern "C" {

void orc_default_config(sdv_config *c) {
    std::memset(c, 0, sizeof(*c));
    c->abi_version = SDV_ABI_VERSION;
    c->device = 0;
    c->max_num_iterations = 20;            // AOptimizer.cpp:380
    c->max_consecutive_invalid_steps = 5;  // ceres default
    c->jacobi_scaling = 1;                 // ceres default
    c->function_tolerance = 1e-3;          // AOptimizer.cpp:384
    c->gradient_tolerance = 1e-10;
    c->parameter_tolerance = 1e-8;
    c->initial_trust_region_radius = 1e4;
    c->max_trust_region_radius = 1e16;
    c->min_trust_region_radius = 1e-32;
    c->min_lm_diagonal = 1e-6;
    c->max_lm_diagonal = 1e32;
    c->min_relative_decrease = 1e-3;
}

// mode 0 = Schur elimination of landmarks, 1 = dense full normal equations (small problems).
int orc_solve_window(const sdv_window *w, const sdv_config *cfg, sdv_delta *out, sdv_stats *st, int mode, int nthreads) {
    auto t0 = std::chrono::steady_clock::now();
    int rc = solve_lm(w, cfg, out, st, mode, nthreads, nullptr, nullptr);
    auto t1 = std::chrono::steady_clock::now();
    st->ms_total_host = std::chrono::duration<double, std::milli>(t1 - t0).count();
    return rc;
}

// AOptimizer::VIInit's solve (AOptimizer.cpp:448-529) on the frames / IMU pairs of `w`: dv[F][3] = the velocity blocks,
// extra[3] = (r_wi_par[0], r_wi_par[1], lambda).  50 iterations (:449, :521) unless the window overrides it; dense normal equations.
int orc_viinit(const sdv_window *w, const sdv_config *cfg, int optim_scale, double *dv, double *extra, sdv_stats *st) {
    auto t0 = std::chrono::steady_clock::now();
    sdv_window ww = *w;
    if (ww.max_num_iterations <= 0) ww.max_num_iterations = 50;
    std::vector<double> dpose(6 * (size_t)w->n_frames), dba(3 * (size_t)w->n_frames), dbg(3 * (size_t)w->n_frames), dl(3 * (size_t)w->n_lmks + 3);
    sdv_delta d;
    d.dpose = dpose.data();
    d.dv = dv;
    d.dba = dba.data();
    d.dbg = dbg.data();
    d.dlmk = dl.data();
    int rc = solve_lm(&ww, cfg, &d, st, 1, 1, nullptr, nullptr, optim_scale == 2 ? 3 : (optim_scale ? 2 : 1), extra); // optim_scale = 2: every block free
    st->ms_total_host = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

// One IMUFactorInit::Evaluate (residuals.hpp:302-410): par = (w_x, w_y, dv_i[3], dv_j[3], dba[3], dbg[3], lambda), pre as in
// orc_imu_factor_eval; r[9], J[9][15] row-major over the concatenated parameter vector.
int orc_imu_init_eval(const double *T_i, const double *T_j, const double *v_i, const double *v_j, double dt, const double *pre,
                      const double *par, double *r, double *J) {
    ImuFactorInit e;
    e.T_fi_w = Aff::From(T_i);
    e.T_fj_w = Aff::From(T_j);
    e.v_i_base = V3::From(v_i);
    e.v_j_base = V3::From(v_j);
    e.dtij = dt;
    e.delta_R = M3::From(pre);
    e.delta_v = V3::From(pre + 9);
    e.delta_p = V3::From(pre + 12);
    e.cov = Mat<9, 9>::From(pre + 15);
    e.J_dR_bg = M3::From(pre + 96);
    e.J_dv_ba = M3::From(pre + 105);
    e.J_dv_bg = M3::From(pre + 114);
    e.J_dp_ba = M3::From(pre + 123);
    e.J_dp_bg = M3::From(pre + 132);
    const int sz[6] = {2, 3, 3, 3, 3, 1}, off[6] = {0, 2, 5, 8, 11, 14};
    const double *pp[6];
    double jb[6][27];
    double *jp[6];
    for (int k = 0; k < 6; k++) {
        pp[k] = par + off[k];
        jp[k] = jb[k];
    }
    if (!e.Evaluate(pp, r, J ? jp : nullptr)) return 1;
    if (J)
        for (int k = 0; k < 6; k++)
            for (int q = 0; q < 9; q++)
                for (int c = 0; c < sz[k]; c++) J[q * 15 + off[k] + c] = jb[k][q * sz[k] + c];
    return 0;
}

// Evaluate every visual residual block at x (NULL = 0): r[O][2], Jp[O][12], Jl[O][6]; returns 1/2 sum r^2 in *cost.
int orc_eval_visual(const sdv_window *w, const sdv_delta *xd, double *r, double *Jp, double *Jl, double *cost) {
    Problem P;
    if (!build_problem(w, P, true)) return SDV_ERR_NUMERICAL_FAILURE;
    std::vector<double> x;
    fill_x(P, xd, x);
    double c = 0;
    for (auto &rb : P.rbs) {
        if (rb.kind != K_VISUAL) continue;
        double rr[2], J[18];
        eval_block(P, rb, x.data(), rr, (Jp || Jl) ? J : nullptr);
        int o = rb.idx;
        if (r) {
            r[2 * o] = rr[0];
            r[2 * o + 1] = rr[1];
        }
        if (Jp) std::memcpy(Jp + 12 * o, J, 12 * sizeof(double));
        if (Jl) std::memcpy(Jl + 6 * o, J + 12, 6 * sizeof(double));
        c += rr[0] * rr[0] + rr[1] * rr[1];
    }
    if (cost) *cost = 0.5 * c;
    return SDV_OK;
}

// Evaluate IMU + bias factors at x: r_imu[P][9], J_imu[P][9*24] (row-major 9x24: pose_i6 pose_j6 vi vj ba bg),
// r_bias[P][6].
int orc_eval_imu(const sdv_window *w, const sdv_delta *xd, double *r_imu, double *J_imu, double *r_bias) {
    Problem P;
    if (!build_problem(w, P, true)) return SDV_ERR_NUMERICAL_FAILURE;
    std::vector<double> x;
    fill_x(P, xd, x);
    for (auto &rb : P.rbs) {
        if (rb.kind == K_IMU) {
            double rr[9], J[9 * 24];
            eval_block(P, rb, x.data(), rr, J);
            std::memcpy(r_imu + 9 * rb.idx, rr, sizeof(rr));
            if (J_imu) {
                // blocks are stored one after the other (9x6, 9x6, 9x3 ...) -> interleave into 9x24 row-major
                const int sz[6] = {6, 6, 3, 3, 3, 3};
                int col = 0;
                size_t off = 0;
                for (int b = 0; b < 6; b++) {
                    for (int i = 0; i < 9; i++)
                        for (int j = 0; j < sz[b]; j++) J_imu[(size_t)rb.idx * 216 + i * 24 + col + j] = J[off + i * sz[b] + j];
                    off += 9 * sz[b];
                    col += sz[b];
                }
            }
        } else if (rb.kind == K_IMU_BIAS && r_bias) {
            double rr[6];
            eval_block(P, rb, x.data(), rr, nullptr);
            std::memcpy(r_bias + 6 * rb.idx, rr, sizeof(rr));
        }
    }
    return SDV_OK;
}

// Total cost (active blocks) and fixed cost at x.
int orc_cost(const sdv_window *w, const sdv_delta *xd, double *cost, double *fixed_cost) {
    Problem P;
    if (!build_problem(w, P, true)) return SDV_ERR_NUMERICAL_FAILURE;
    std::vector<double> x;
    fill_x(P, xd, x);
    std::vector<double> res(P.nres);
    evaluate(P, x.data(), cost, res.data(), nullptr, 1);
    if (fixed_cost) *fixed_cost = fixed_cost_of(P, x.data());
    return SDV_OK;
}

// Reduced (Schur) system at x = 0 restricted to the landmarks [l0, l1) (+ the non-visual factors when with_factors):
//   S = sum_{obs of those landmarks} Jp^T Jp - sum_l W_l (H_ll + lambda I)^-1 W_l^T  [+ factor J^T J],   g likewise.
// Used by the multi-process (gloo) test of the landmark sharding: the partial systems of the ranks must add up to the full one.
// S is n x n row-major (both triangles), columns in the oracle's dense-block order; returns n in *n_out.
int orc_reduced_system(const sdv_window *w, int l0, int l1, int with_factors, double lambda, double *S, double *g, int *n_out) {
    Problem P;
    if (!build_problem(w, P, true)) return SDV_ERR_NUMERICAL_FAILURE;
    const int n = P.ndense;
    if (n_out) *n_out = n;
    if (!S || !g) return SDV_OK;
    std::vector<double> x(P.nx, 0.0), res(P.nres), jac(P.jsize);
    double cost;
    evaluate(P, x.data(), &cost, res.data(), jac.data(), 1);
    for (size_t i = 0; i < (size_t)n * n; i++) S[i] = 0;
    for (int i = 0; i < n; i++) g[i] = 0;
    struct LAcc { M3 V; V3 gl; std::vector<std::pair<int, std::array<double, 18>>> W; bool used = false; };
    std::vector<LAcc> la(P.L);
    for (auto &a : la) { a.V = M3::Zero(); a.gl = V3::Zero(); }
    for (auto &rb : P.rbs) {
        if (!rb.active) continue;
        int lm = -1;
        for (int id : rb.pb)
            if (id >= P.lmk_id(0) && P.pbs[id].elim) lm = id - P.lmk_id(0);
        if (lm < 0 && !with_factors) continue;
        if (lm >= 0 && (lm < l0 || lm >= l1)) continue;
        size_t o1 = 0;
        const double *Jl = nullptr;
        {
            size_t o = 0;
            for (int a = 0; a < rb.npb; a++) {
                if (lm >= 0 && rb.pb[a] == P.lmk_id(lm)) Jl = jac.data() + rb.joff + o;
                o += (size_t)rb.nres * P.pbs[rb.pb[a]].size;
            }
        }
        const double *r = res.data() + rb.roff;
        if (Jl) {
            LAcc &a = la[lm];
            a.used = true;
            for (int i = 0; i < 3; i++) {
                for (int j = 0; j < 3; j++)
                    for (int q = 0; q < rb.nres; q++) a.V(i, j) += Jl[q * 3 + i] * Jl[q * 3 + j];
                for (int q = 0; q < rb.nres; q++) a.gl[i] += Jl[q * 3 + i] * r[q];
            }
        }
        for (int a = 0; a < rb.npb; a++) {
            const PBlock &pa = P.pbs[rb.pb[a]];
            const double *Ja = jac.data() + rb.joff + o1;
            o1 += (size_t)rb.nres * pa.size;
            if (!pa.active || pa.elim) continue;
            for (int q = 0; q < rb.nres; q++)
                for (int i = 0; i < pa.size; i++) g[pa.col + i] += Ja[q * pa.size + i] * r[q];
            size_t o2 = 0;
            for (int b = 0; b < rb.npb; b++) {
                const PBlock &pb = P.pbs[rb.pb[b]];
                const double *Jb = jac.data() + rb.joff + o2;
                o2 += (size_t)rb.nres * pb.size;
                if (!pb.active || pb.elim) continue;
                for (int i = 0; i < pa.size; i++)
                    for (int j = 0; j < pb.size; j++) {
                        double sacc = 0;
                        for (int q = 0; q < rb.nres; q++) sacc += Ja[q * pa.size + i] * Jb[q * pb.size + j];
                        S[(size_t)(pa.col + i) * n + pb.col + j] += sacc;
                    }
            }
            if (Jl) {
                LAcc &acc = la[lm];
                std::array<double, 18> *Wp = nullptr;
                for (auto &e : acc.W)
                    if (e.first == pa.col) Wp = &e.second;
                if (!Wp) {
                    acc.W.push_back({pa.col, std::array<double, 18>{}});
                    Wp = &acc.W.back().second;
                }
                for (int i = 0; i < pa.size; i++)
                    for (int j = 0; j < 3; j++)
                        for (int q = 0; q < rb.nres; q++) (*Wp)[i * 3 + j] += Ja[q * pa.size + i] * Jl[q * 3 + j];
            }
        }
    }
    for (int l = l0; l < l1 && l < P.L; l++) {
        LAcc &a = la[l];
        if (!a.used) continue;
        M3 V = a.V;
        for (int i = 0; i < 3; i++) V(i, i) += lambda;
        M3 Vi = inverse3(V);
        for (auto &ea : a.W) {
            double Y[18];
            for (int i = 0; i < 6; i++)
                for (int j = 0; j < 3; j++) {
                    double sacc = 0;
                    for (int q = 0; q < 3; q++) sacc += ea.second[i * 3 + q] * Vi(q, j);
                    Y[i * 3 + j] = sacc;
                }
            for (int i = 0; i < 6; i++) {
                double sacc = 0;
                for (int q = 0; q < 3; q++) sacc += Y[i * 3 + q] * a.gl[q];
                g[ea.first + i] -= sacc;
            }
            for (auto &eb : a.W)
                for (int i = 0; i < 6; i++)
                    for (int j = 0; j < 6; j++) {
                        double sacc = 0;
                        for (int q = 0; q < 3; q++) sacc += Y[i * 3 + q] * eb.second[j * 3 + q];
                        S[(size_t)(ea.first + i) * n + eb.first + j] -= sacc;
                    }
        }
    }
    return SDV_OK;
}

// Generic single-functor entry points for the known-answer tests -----------------------------------------------
int orc_exp_so3(const double *v, double *R) { exp_so3(V3::From(v)).to(R); return 0; }
int orc_log_so3(const double *R, double *v) { log_so3(M3::From(R)).to(v); return 0; }
int orc_right_jacobian(const double *v, double *J) { so3_rightJacobian(V3::From(v)).to(J); return 0; }

int orc_angular_eval(const double *bearing, const double *T_s_f, const double *T_f_w, const double *t_w_lmk, double sigma,
                     const double *dx6, const double *dp3, double *r2, double *J26, double *J23) {
    AngularErr e{V3::From(bearing), Aff::From(T_s_f), Aff::From(T_f_w), V3::From(t_w_lmk), sigma};
    const double *par[2] = {dx6, dp3};
    double *J[2] = {J26, J23};
    return e.Evaluate(par, r2, (J26 || J23) ? J : nullptr) ? 0 : 1;
}
int orc_reproj_eval(const double *uv, const double *K4, const double *T_s_f, const double *T_f_w, const double *t_w_lmk,
                    double sigma, const double *dx6, const double *dp3, double *r2, double *J26, double *J23) {
    ReprojErr e;
    e.p2d[0] = uv[0];
    e.p2d[1] = uv[1];
    for (int k = 0; k < 4; k++) e.K[k] = K4[k];
    e.T_s_f = Aff::From(T_s_f);
    e.T_f_w_base = Aff::From(T_f_w);
    e.t_w_lmk = V3::From(t_w_lmk);
    e.sigma = sigma;
    const double *par[2] = {dx6, dp3};
    double *J[2] = {J26, J23};
    return e.Evaluate(par, r2, (J26 || J23) ? J : nullptr) ? 0 : 1;
}
int orc_pose_prior_eval(const double *T, const double *T_prior, const double *sqrt_inf_diag6, const double *dx6, double *r6,
                        double *J66) {
    PosePrior e;
    e.T = Aff::From(T);
    e.T_prior = Aff::From(T_prior);
    e.sqrt_inf = Mat<6, 6>::Zero();
    for (int k = 0; k < 6; k++) e.sqrt_inf(k, k) = sqrt_inf_diag6[k];
    const double *par[1] = {dx6};
    double *J[1] = {J66};
    return e.Evaluate(par, r6, J66 ? J : nullptr) ? 0 : 1;
}
int orc_p2l_eval(const double *delta, const double *T_f_w, const double *t_w_lmk, const double *sqrt_inf9, const double *dx6,
                 const double *dp3, double *r3, double *J36, double *J33) {
    PoseToLandmark e{V3::From(delta), V3::From(t_w_lmk), Aff::From(T_f_w), M3::From(sqrt_inf9)};
    const double *par[2] = {dx6, dp3};
    double *J[2] = {J36, J33};
    return e.Evaluate(par, r3, (J36 || J33) ? J : nullptr) ? 0 : 1;
}
// IMU factor with explicit inputs: pre = [dR9 dv3 dp3 cov81 J_dR_bg9 J_dv_ba9 J_dv_bg9 J_dp_ba9 J_dp_bg9] (141 doubles);
// params = [dTi6 dTj6 dvi3 dvj3 dba3 dbg3] (24); outputs r9 and J (9x24 row-major, may be NULL).
int orc_imu_factor_eval(const double *T_i, const double *T_j, const double *v_i, const double *v_j, double dt, const double *pre,
                        const double *params24, double *r9, double *J924) {
    ImuFactor e;
    e.T_fi_w_base = Aff::From(T_i);
    e.T_fj_w_base = Aff::From(T_j);
    e.v_i_base = V3::From(v_i);
    e.v_j_base = V3::From(v_j);
    e.dtij = dt;
    e.delta_R = M3::From(pre);
    e.delta_v = V3::From(pre + 9);
    e.delta_p = V3::From(pre + 12);
    e.cov = Mat<9, 9>::From(pre + 15);
    e.J_dR_bg = M3::From(pre + 96);
    e.J_dv_ba = M3::From(pre + 105);
    e.J_dv_bg = M3::From(pre + 114);
    e.J_dp_ba = M3::From(pre + 123);
    e.J_dp_bg = M3::From(pre + 132);
    const double *par[6] = {params24, params24 + 6, params24 + 12, params24 + 15, params24 + 18, params24 + 21};
    double Jb[9 * 24];
    double *J[6] = {Jb, Jb + 54, Jb + 108, Jb + 135, Jb + 162, Jb + 189};
    if (!e.Evaluate(par, r9, J924 ? J : nullptr)) return 1;
    if (J924) {
        const int sz[6] = {6, 6, 3, 3, 3, 3};
        int col = 0;
        size_t off = 0;
        for (int b = 0; b < 6; b++) {
            for (int i = 0; i < 9; i++)
                for (int j = 0; j < sz[b]; j++) J924[i * 24 + col + j] = Jb[off + i * sz[b] + j];
            off += 9 * sz[b];
            col += sz[b];
        }
    }
    return 0;
}

// IMU pre-integration state, flattened for ctypes:
//  [acc3 gyr3 ba3 bg3 v3 T_f_w12 is_kf1 dR9 dv3 dp3 Sigma81 J_dR_bg9 J_dv_ba9 J_dv_bg9 J_dp_ba9 J_dp_bg9] = 169 doubles
#define ORC_IMU_STATE_SIZE 169
static void imu_unpack(const double *p, ImuState &s) {
    s.acc = V3::From(p);
    s.gyr = V3::From(p + 3);
    s.ba = V3::From(p + 6);
    s.bg = V3::From(p + 9);
    s.v = V3::From(p + 12);
    s.T_f_w = Aff::From(p + 15);
    s.frame_is_kf = p[27] != 0.0;
    s.delta_R = M3::From(p + 28);
    s.delta_v = V3::From(p + 37);
    s.delta_p = V3::From(p + 40);
    s.Sigma = Mat<9, 9>::From(p + 43);
    s.J_dR_bg = M3::From(p + 124);
    s.J_dv_ba = M3::From(p + 133);
    s.J_dv_bg = M3::From(p + 142);
    s.J_dp_ba = M3::From(p + 151);
    s.J_dp_bg = M3::From(p + 160);
}
static void imu_pack(const ImuState &s, double *p) {
    s.acc.to(p);
    s.gyr.to(p + 3);
    s.ba.to(p + 6);
    s.bg.to(p + 9);
    s.v.to(p + 12);
    s.T_f_w.to(p + 15);
    p[27] = s.frame_is_kf ? 1.0 : 0.0;
    s.delta_R.to(p + 28);
    s.delta_v.to(p + 37);
    s.delta_p.to(p + 40);
    s.Sigma.to(p + 43);
    s.J_dR_bg.to(p + 124);
    s.J_dv_ba.to(p + 133);
    s.J_dv_bg.to(p + 142);
    s.J_dp_ba.to(p + 151);
    s.J_dp_bg.to(p + 160);
}
int orc_imu_state_size(void) { return ORC_IMU_STATE_SIZE; }
int orc_process_imu(const double *last, const double *kf_ba, const double *kf_bg, double dt, const double *eta6, double rate_hz,
                    double *cur /* in: acc,gyr,is_kf ; out: everything */) {
    ImuState l, c;
    imu_unpack(last, l);
    imu_unpack(cur, c);
    process_imu(l, V3::From(kf_ba), V3::From(kf_bg), dt, eta6, rate_hz, c);
    imu_pack(c, cur);
    return 0;
}
int orc_bias_delta_correction(double *state, const double *d_ba, const double *d_bg) {
    ImuState s;
    imu_unpack(state, s);
    bias_delta_correction(s, V3::From(d_ba), V3::From(d_bg));
    imu_pack(s, state);
    return 0;
}

int orc_abi_version(void) { return SDV_ABI_VERSION; }
}

extern "C" {
// Symmetric eigen-decomposition used by the marginal-prior construction: method 0 = Householder + implicit QL (what
// schur_prior uses), 1 = cyclic Jacobi (independent cross-check).  w [n] ascending, V [n*n] row-major, columns = eigenvectors.
int orc_sym_eig(int n, const double *A, int method, double *w, double *V) {
    std::vector<double> vw, vV;
    if (method == 0) orc::sym_eig(n, A, vw, vV);
    else orc::jacobi_eig(n, A, vw, vV);
    std::copy(vw.begin(), vw.end(), w);
    std::copy(vV.begin(), vV.end(), V);
    return 0;
}
// Dense core of the marginal-prior construction (marg.hpp).  Returns 1 on success, 0 when the reference returns false.
// Output buffers must hold n*n, n, n*n, n, n*n, n doubles; *n_full receives the rank.
int orc_schur_prior(int m, int n, const double *A, const double *b, double eps, double *Ak, double *bk, int *n_full, double *U, double *Lambda,
                    double *Jm, double *r0) {
    std::vector<double> vAk, vbk, vU, vL, vJ, vr;
    int nf = 0;
    if (!orc::schur_prior(m, n, A, b, eps, vAk, vbk, nf, vU, vL, vJ, vr)) return 0;
    std::copy(vAk.begin(), vAk.end(), Ak);
    std::copy(vbk.begin(), vbk.end(), bk);
    std::copy(vU.begin(), vU.end(), U);
    std::copy(vL.begin(), vL.end(), Lambda);
    std::copy(vJ.begin(), vJ.end(), Jm);
    std::copy(vr.begin(), vr.end(), r0);
    *n_full = nf;
    return 1;
}
}

// Device-visible problem description and solver state shared by all kernels.
#pragma once
#include <stdint.h>

namespace sdv {

// Row of the frame-camera table (doubles): everything a visual factor needs about its (frame, camera) pair at the
// current linearisation point. 32 doubles = 256 B so rows stay 16-byte aligned for cp.async.bulk staging.
//   [0..8]   Rsw = R_s_f * R_f_w * exp(dw)            [9..11]  tsw = R_s_f (R_f_w dt + t_f_w) + t_s_f
//   [12..20] G   = R_s_f * R_f_w  (base, constant)     [21..29] Jq  (angular: Jr(log exp(dw)); pixel: Jr(log R')Jr(log R')^-1 Jr(dw))
//   [30] default weight 1/sigma of the camera          [31] unused
// Row stride of the frame-camera table (32 used + 2 pad doubles): 16-byte aligned, and lanes of a warp that touch up to 8
// different rows at the same offset hit different shared-memory banks.  Global and shared copies use the same layout so the
// whole table moves with a handful of large TMA bulk copies.
constexpr int FCT_ROW = 34;

// Per-solve accumulators (zeroed by the control kernels).
struct Accum {
    double cost[2];        // 1/2 sum r^2 of linearisation buffer 0 / 1
    double model_gd;       // sum g.delta   (landmark part + pose part)
    double model_dd;       // sum d_i delta_i^2
    double step_norm2;     // |delta|^2
    double cand_norm2;     // |x + delta|^2
    double fixed_cost;     // 1/2 sum r^2 of residual blocks whose parameter blocks are all constant (Ceres fixed_cost)
    unsigned long long grad_max_bits; // max |g| over landmark columns (bit pattern of a non-negative double)
    int schur_fail;        // a landmark block was not positive definite
    int chol_fail;         // reduced system not positive definite
};

struct LMState {
    int iter;              // trust-region steps attempted so far (Ceres iteration counter)
    int status;            // 0 = running, otherwise 1 + sdv_termination
    int cur;               // linearisation buffer holding J(x)
    int step_valid;        // the step computed this iteration is valid (model decrease > 0, factorisation ok)
    int have_cand;         // candidate evaluated
    int num_consecutive_invalid;
    int atleast_one_successful;
    int n_ok, n_bad;
    int scaling_done;      // jacobi scaling vector computed (iteration 0)
    int need_grad_check;   // gradient tolerance test pending (iteration 0 and after each accepted step)
    int pad0;
    double radius, decrease_factor;
    double x_cost, cand_cost, model_cost_change;
    double x_norm2;
    double initial_cost;
    double trace_cost[64];
    double trace_radius[64];
    double trace_model[64];
    int trace_accepted[64];
};

struct SolverOpts {
    int max_num_iterations, max_consecutive_invalid_steps, jacobi_scaling;
    double function_tolerance, gradient_tolerance, parameter_tolerance;
    double initial_radius, max_radius, min_radius, min_diag, max_diag, min_relative_decrease;
};

struct DevProblem {
    // sizes
    int F, C, L, O, P;
    int vio, kind;
    int n;        // reduced system dimension
    int n_pad;    // n rounded up to a multiple of 32
    int ld;       // leading dimension of S (= n_pad)
    int band_bw;  // > 0: half-bandwidth of S in 16-column blocks, the banded single-CTA Cholesky (k_chol_band) is used
    int nslots;   // number of (landmark, frame) slots of this rank
    int l0, l1;   // landmark range owned by this rank
    int o0, o1;   // observation range owned by this rank
    int rank, world;
    int fct_in_smem; // stage the frame-camera table in shared memory
    int Ocap;        // stride of the r / J planes of this rank: real observations + 2 pseudo-observations per PoseToLandmark factor
    // the other AOptimizer solves as masks of the window solve (sdv_window::visual_loss_huber_a / landmarks_constant / max_num_iterations)
    int lmk_const;   // every landmark block is constant (single-frame solves): no elimination, no landmark update
    int max_iter;    // > 0: iteration cap of this window (overrides SolverOpts::max_num_iterations)
    double huber_a;  // > 0: ceres::HuberLoss(a) + Corrector on every visual residual block
    // device structure pass (sdv_struct.cuh): scratch of the scans and the bucket size of the tiling; st_on = 0: the host built the lists
    int *st_head, *st_sidx, *st_ssum, *st_thead, *st_tidx, *st_tsum, *st_tot;
    int st_capq, st_on;
    // frames
    const double *T_f_w, *v, *ba, *bg;
    const unsigned char *has_prior;
    const double *T_prior, *inf_prior;
    const int *pose_col, *vb_col; // column of pose(6) / v,ba,bg(9) in the reduced system, -1 = constant or unused
    // cameras
    const double *T_s_f, *K, *cam_w;
    // landmarks
    const double *lmk_t;
    const int *lmk_col;     // column in the reduced system for dense (kept) landmarks, -1 = eliminated
    const uint32_t *tile_nz; // [(n_pad/32 + 1)][4] bit j of row i: tile (i, j) of the Cholesky factor is structurally non-zero
    const int *slot_ptr;    // [L+1] first slot of each landmark
    const int *tile_ptr;    // [ntiles+1] landmark ranges of this rank with <= FT slots and <= FT_LMK landmarks (k_lin_schur, k_backsub_cost)
    int ntiles;
    const int *slot_frame;  // [nslots]
    const int *slot_obs_ptr;// [nslots+1]
    const int *slot_obs;    // plane indices (local to this rank) of the observations grouped by slot
    // observations (indices relative to the full window; buffers of this rank are indexed o - o0)
    const int *obs_lmk, *obs_fc;
    const double *obs_meas; // [O][3] bearing or [O][2] uv (array-of-structs as the caller provides it)
    const double *obs_w;    // per-observation 1/sigma or nullptr
    // imu
    const int *imu_i, *imu_j;
    const double *imu_dt, *imu_dR, *imu_dv, *imu_dp, *imu_cov;
    const double *imu_J_dR_bg, *imu_J_dv_ba, *imu_J_dv_bg, *imu_J_dp_ba, *imu_J_dp_bg;
    const double *imu_sigma_ba, *imu_sigma_bg;
    double *imu_inf_sqrt;   // [P][81] upper-triangular sqrt information, computed once per upload
    // dense prior
    int mp_nfull, mp_n, mp_nblk; // rows, cols of J_m; number of 3-column groups mapped
    const double *mp_J, *mp_r0;
    const int *mp_src_col;  // [mp_nmap] column in J_m
    const int *mp_dst_col;  // [mp_nmap] column in the reduced system (-1 = constant)
    int mp_nmap;
    double *mp_H;           // [mp_nmap][mp_nmap] J_m^T J_m restricted to mapped columns
    double *mp_g0;          // [mp_nmap] J_m^T r0
    // sparsified prior (AngularAdjustmentCERESAnalytic.cpp:387-483)
    int sp_has_imu, sp_frame, sp_has_lmk, sp_lmk0, sp_np2l, sp_nl2l;
    const double *sp_blob;     // [T_prior 12 | v_prior 3 | ba_prior 3 | bg_prior 3 | imu_sqrt_inf 225 | lmk_prior 3 | lmk_sqrt_inf 9]
    const int *sp_p2l_lmk;     // [np2l]
    const int *sp_p2l_plane;   // [np2l] first of the two pseudo-observation plane indices, -1 = landmark owned by another rank
    const double *sp_p2l_delta, *sp_p2l_sqrt; // [np2l][3], [np2l][9]
    const int *sp_l2l_a, *sp_l2l_b;
    const double *sp_l2l_delta, *sp_l2l_sqrt;
};

// One linearisation (residuals + Jacobians at a point). Two of these are alive: x and the candidate.
struct LinBuf {
    double *fct;     // [F*C][FCT_ROW]
    double *r;       // [2][Oloc]
    double *Jp;      // [12][Oloc]
    double *Jl;      // [6][Oloc]
    double *imu_r;   // [P][9]
    double *imu_J;   // [P][9*24]
    double *bias_r;  // [P][6]
    double *prior_r; // [F][6]
    double *prior_J; // [F][36]
    double *mp_r;    // [mp_nfull]
    double *sp_r;    // [15 + 3 + 3*nl2l] IMUPriordx, Landmark3DPrior, LandmarkToLandmark residuals
    double *sp_J;    // [15*15] IMUPriordx Jacobian (pose6 | v3 | ba3 | bg3)
    double *xp;      // [n_pad] reduced parameters (pose6,v3,ba3,bg3 per frame ... dense landmarks)
    double *xl;      // [3L] eliminated landmark parameters
};

} // namespace sdv

/*
 * sdv.h — C ABI of the B200-native sliding-window BA/VIO backend.
 *
 * This is the drop-in boundary for the one hot path this repository replaces:
 * SaDVIO's windowed optimisation `isae::AOptimizer::localMapVIOptimization`
 * (reference cpp/src/optimizers/AOptimizer.cpp:352-446) and
 * `isae::AOptimizer::localMapBA` (AOptimizer.cpp:299-350), i.e. the Ceres
 * problem build + `ceres::Solve` + the cost functors it calls back into.
 *
 * The reference has no FFI: the boundary upstream is the C++ abstract class
 * `isae::AOptimizer` (cpp/include/isaeslam/optimizers/AOptimizer.h:13-91).
 * Every entry point below names the reference code it replaces.  Signatures
 * are plain C: pointers, sizes, POD structs — no torch / Eigen / STL types.
 * All floating point data is FP64 (the reference is `double` throughout),
 * indices are int32, matrices are row-major, rigid transforms are 3x4
 * row-major [R | t] (the top three rows of Eigen::Affine3d::matrix()).
 *
 * Ownership: the caller owns every buffer reachable from `sdv_window`,
 * `sdv_delta` and `sdv_stats`; the library copies host->device inside the
 * call and keeps nothing but device scratch tied to the handle.  A handle is
 * bound to one CUDA device, is not re-entrant, and mirrors one
 * `isae::AOptimizer` instance (the back-end one, slamParameters.cpp:273-274).
 *
 * Errors: every function returns an `sdv_status` (0 = ok).  No exception
 * crosses the ABI.  The reference returns `bool` and practically always
 * `true`; the adapters (sadvio_b200/host/b200_optimizer.hpp, sadvio_b200/api.py)
 * map a status that means "no solve ran" to `false` and leave the state
 * untouched; SDV_ERR_NUMERICAL_FAILURE (Ceres FAILURE) still carries the last
 * accepted x, which they write back, returning `true` as the reference does.
 */
#ifndef SDV_H
#define SDV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDV_ABI_VERSION 2
#define SDV_MAX_TRACE 64

typedef enum sdv_status {
    SDV_OK = 0,
    SDV_ERR_INVALID_ARGUMENT = 1, /* null pointer, negative size, index out of range */
    SDV_ERR_CUDA = 2,             /* a CUDA runtime call failed (see sdv_last_error) */
    SDV_ERR_NO_DEVICE = 3,        /* no usable sm_100 device: there is NO CPU fallback */
    SDV_ERR_UNSUPPORTED = 4,      /* structurally valid input this build cannot solve */
    SDV_ERR_NUMERICAL_FAILURE = 5,/* Ceres would report FAILURE (5 invalid steps in a row) */
    SDV_ERR_COMM = 6              /* NCCL communicator problem */
} sdv_status;

/* Which visual cost functor the window uses (one optimizer class = one kind). */
typedef enum sdv_factor_kind {
    /* AngularAdjustmentCERESAnalytic::AngularErrCeres_pointxd_dx
       (AngularAdjustmentCERESAnalytic.h:45-120) — the default optimizer. */
    SDV_FACTOR_ANGULAR = 0,
    /* BundleAdjustmentCERESAnalytic::ReprojectionErrCeres_pointxd_dx
       (BundleAdjustmentCERESAnalytic.h:41-98) + Camera::project (Camera.cpp:84-139). */
    SDV_FACTOR_PIXEL = 1
} sdv_factor_kind;

/* Termination, named after ceres::TerminationType + the message it carries. */
typedef enum sdv_termination {
    SDV_TERM_NO_CONVERGENCE = 0,      /* max_num_iterations reached */
    SDV_TERM_FUNCTION_TOLERANCE = 1,  /* |dcost| <= function_tolerance * cost (candidate NOT applied) */
    SDV_TERM_GRADIENT_TOLERANCE = 2,  /* max|g| <= gradient_tolerance */
    SDV_TERM_PARAMETER_TOLERANCE = 3, /* |step| <= parameter_tolerance * (|x| + parameter_tolerance) */
    SDV_TERM_MIN_RADIUS = 4,          /* trust region radius fell below min_trust_region_radius */
    SDV_TERM_FAILURE = 5              /* max_num_consecutive_invalid_steps reached */
} sdv_termination;

/*
 * Solver options.  Defaults (sdv_default_config) are the options the reference
 * sets at AOptimizer.cpp:376-388 plus the Ceres 2.2.0 defaults it leaves
 * untouched (docker/Dockerfile:50 pins 2.2.0; Ceres itself is not part of
 * /root/reference — see oracle/README.md "parity pin").
 */
typedef struct sdv_config {
    int32_t abi_version;             /* SDV_ABI_VERSION */
    int32_t device;                  /* CUDA device ordinal */
    int32_t max_num_iterations;      /* 20     AOptimizer.cpp:380 */
    int32_t max_consecutive_invalid_steps; /* 5   Ceres default */
    int32_t jacobi_scaling;          /* 1      Ceres default */
    int32_t reserved0;
    double function_tolerance;       /* 1e-3   AOptimizer.cpp:384 */
    double gradient_tolerance;       /* 1e-10  Ceres default */
    double parameter_tolerance;      /* 1e-8   Ceres default */
    double initial_trust_region_radius; /* 1e4 */
    double max_trust_region_radius;  /* 1e16 */
    double min_trust_region_radius;  /* 1e-32 */
    double min_lm_diagonal;          /* 1e-6 */
    double max_lm_diagonal;          /* 1e32 */
    double min_relative_decrease;    /* 1e-3 */
} sdv_config;

/*
 * Dense marginalisation prior = isae::MarginalizationFactor
 * (marginalization.hpp:88-218): r = r0 + J * dx, J is n_full x n, dx gathers
 * the kept frame's (pose6, v3, ba3, bg3) at column `frame_col` and each kept
 * landmark's dp3 at `keep_col[k]` (keep_col < 0 = landmark skipped, as
 * `_map_lmk_idx == -1` at marginalization.hpp:139).  Kept landmarks couple
 * to each other through J, so they live in the reduced (dense) system,
 * mirroring elimination group 2 (AngularAdjustmentCERESAnalytic.cpp:352-379).
 */
typedef struct sdv_dense_prior {
    int32_t n_full;          /* rows of J  (_n_full) */
    int32_t n;               /* cols of J  (_n) */
    const double *J;         /* [n_full][n] row-major  (_marginalization_jacobian) */
    const double *r0;        /* [n_full]               (_marginalization_residual) */
    int32_t frame;           /* index into frames[] of _frame_to_keep, or -1 */
    int32_t frame_col;       /* _map_frame_idx.at(_frame_to_keep) */
    int32_t n_keep;          /* number of kept landmarks */
    int32_t reserved0;
    const int32_t *keep_lmk; /* [n_keep] landmark indices, _lmk_to_keep iteration order */
    const int32_t *keep_col; /* [n_keep] _map_lmk_idx.at(lmk) */
} sdv_dense_prior;

/*
 * Sparsified prior (AngularAdjustmentCERESAnalytic.cpp:387-483):
 *  VIO case: one IMUPriordx (residuals.hpp:634-700) on the kept frame and one
 *            PoseToLandmarkFactor (residuals.hpp:561-599) per kept landmark;
 *  VO case : one Landmark3DPrior (residuals.hpp:506-526) and a chain of
 *            LandmarkToLandmarkFactor (residuals.hpp:528-559).
 */
typedef struct sdv_sparse_prior {
    /* --- VIO --- */
    int32_t has_imu_prior;      /* 1 => IMUPriordx on `frame` */
    int32_t frame;              /* index into frames[] of frame_to_keep */
    double T_prior[12];         /* _T_prior */
    double v_prior[3], ba_prior[3], bg_prior[3];
    double imu_sqrt_inf[225];   /* 15x15 row-major, _map_frame_inf.at(frame_to_keep) */
    int32_t n_p2l;              /* PoseToLandmarkFactor count */
    int32_t reserved0;
    const int32_t *p2l_lmk;     /* [n_p2l] */
    const double *p2l_delta;    /* [n_p2l][3]  _map_lmk_prior.at(lmk) */
    const double *p2l_sqrt_inf; /* [n_p2l][9]  _map_lmk_inf.at(lmk) */
    /* --- VO --- */
    int32_t has_lmk_prior;      /* 1 => Landmark3DPrior on lmk0 */
    int32_t lmk0;               /* _lmk_with_prior */
    double lmk_prior[3];        /* _prior_lmk */
    double lmk_sqrt_inf[9];     /* _info_lmk */
    int32_t n_l2l;              /* LandmarkToLandmarkFactor count */
    int32_t reserved1;
    const int32_t *l2l_a;       /* [n_l2l] lmk_k   */
    const int32_t *l2l_b;       /* [n_l2l] lmk_kp1 */
    const double *l2l_delta;    /* [n_l2l][3] */
    const double *l2l_sqrt_inf; /* [n_l2l][9] */
} sdv_sparse_prior;

/*
 * One sliding window, flattened to structure-of-arrays in the order the
 * reference walks its pointer graph:
 *   frames    NEWEST -> OLDEST, as LocalMap::getLastNFramesIn fills
 *             frame_vector (amap.h:28-32); the LAST `n_fixed` are constant
 *             (AngularAdjustmentCERESAnalytic.cpp:234-236, AOptimizer.cpp:46-51);
 *   landmarks local_map->getLandmarks()["pointxd"] order, filtered by
 *             isInitialized && !isOutlier (…Analytic.cpp:254);
 *   obs       landmark-major, lmk->getFeatures() order within a landmark,
 *             filtered as at …Analytic.cpp:272-275  => obs_lmk is non-decreasing;
 *   imu       one entry per (framei = lastKF(framej), framej) pair that passes
 *             the tests at AOptimizer.cpp:55-71, in frame_vector order of j.
 */
typedef struct sdv_window {
    int32_t abi_version;     /* SDV_ABI_VERSION */
    int32_t vio;             /* 1: localMapVIOptimization (pose6+v3+ba3+bg3 per frame); 0: localMapBA (pose6) */
    int32_t factor_kind;     /* sdv_factor_kind */
    int32_t n_frames;
    int32_t n_fixed;         /* fixed_frame_number argument */
    int32_t n_cams;
    int32_t n_lmks;
    int32_t n_obs;
    int32_t n_imu;
    int32_t reserved0;           /* 0 */

    /* frames */
    const double *T_f_w;     /* [F][12] Frame::getWorld2FrameTransform() */
    const double *v;         /* [F][3]  IMU::getVelocity()   (vio) */
    const double *ba;        /* [F][3]  IMU::getBa()         (vio) */
    const double *bg;        /* [F][3]  IMU::getBg()         (vio) */
    const uint8_t *has_imu;  /* [F]     frame->getIMU() != nullptr (vio); NULL => all 1.  Informational: a frame
                                        without IMU appears in no imu[] entry (AOptimizer.cpp:60-73), so its v / ba /
                                        bg columns carry no factor and their updates stay exactly 0 — the same
                                        result as not creating the blocks (AOptimizer.cpp:30-52) */
    const uint8_t *has_prior;/* [F]     Frame::hasPrior(); NULL => none */
    const double *T_prior;   /* [F][12] Frame::getPrior()      (used where has_prior) */
    const double *inf_prior; /* [F][6]  Frame::getInfPrior(), used AS sqrt-information
                                        (…Analytic.cpp:241 passes it .asDiagonal() to PosePriordx) */

    /* cameras: one entry per distinct (extrinsics, intrinsics) sensor model */
    const double *T_s_f;     /* [C][12] ASensor::getFrame2SensorTransform() */
    const double *K;         /* [C][4]  fx, fy, cx, cy  (Camera::getCalibration()); focal = (fx+fy)/2 (Camera.h:46) */

    /* landmarks */
    const double *lmk_t;     /* [L][3]  ALandmark::getPose().translation() */

    /* observations (visual residual blocks) */
    const int32_t *obs_lmk;  /* [O] */
    const int32_t *obs_frame;/* [O] */
    const int32_t *obs_cam;  /* [O] */
    const double *obs_bearing; /* [O][3] AFeature::getBearingVectors().at(0)  (SDV_FACTOR_ANGULAR) */
    const double *obs_uv;      /* [O][2] AFeature::getPoints().at(0)          (SDV_FACTOR_PIXEL)   */
    const double *obs_sigma;   /* [O] or NULL => reference default: 1.5/focal (angular, …Analytic.cpp:283), 1.0 (pixel) */

    /* IMU pre-integration factors: IMUFactor + IMUBiasFactor per entry (AOptimizer.cpp:72-93) */
    const int32_t *imu_i;    /* [P] frame index of framei */
    const int32_t *imu_j;    /* [P] frame index of framej */
    const double *imu_dt;    /* [P] (ts_j - ts_i) * 1e-9 */
    const double *imu_dR;    /* [P][9]  imu_j->getDeltaR() */
    const double *imu_dv;    /* [P][3]  getDeltaV() */
    const double *imu_dp;    /* [P][3]  getDeltaP() */
    const double *imu_cov;   /* [P][81] getCov() */
    const double *imu_J_dR_bg; /* [P][9] */
    const double *imu_J_dv_ba; /* [P][9] */
    const double *imu_J_dv_bg; /* [P][9] */
    const double *imu_J_dp_ba; /* [P][9] */
    const double *imu_J_dp_bg; /* [P][9] */
    const double *imu_sigma_ba; /* [P] imu_i->getbAccNoise() (residuals.hpp:259) */
    const double *imu_sigma_bg; /* [P] imu_i->getbGyrNoise() (residuals.hpp:261) */

    /* priors injected by addMarginalizationResiduals (…Analytic.cpp:341-486); either may be NULL */
    const sdv_dense_prior *dense_prior;
    const sdv_sparse_prior *sparse_prior;

    /* --- the other AOptimizer solves as masks of the window solve --- */
    double visual_loss_huber_a;  /* 0: loss_function = nullptr (localMapBA, localMapVIOptimization, singleFrameOptimization);
                                    a > 0: ceres::HuberLoss(a) on every VISUAL residual block — landmarkOptimization
                                    (AOptimizer.cpp:102) and singleFrameVIOptimization (:223) pass sqrt(1.345); the IMU
                                    factors of the latter are added with a null loss (:239) */
    int32_t landmarks_constant;  /* 1: every landmark block is SetParameterBlockConstant — the single-frame solves
                                    (addSingleFrameResiduals, …Analytic.cpp:38-43) */
    int32_t max_num_iterations;  /* > 0 overrides sdv_config::max_num_iterations for this window: 10 in landmarkOptimization
                                    (AOptimizer.cpp:113), 5 in the single-frame solves (:166, :247) */
    const uint8_t *lmk_has_prior;/* [L] or NULL — read by sdv_marginalize only: ALandmark::hasPrior(), the sticky flag earlier
                                    marginalisations set on the landmarks they kept (marginalization.cpp:72-79,
                                    ALandmark.h:100-107); NULL = the landmarks kept by `dense_prior` */
} sdv_window;
/*
 * How the four solves map onto sdv_window (the adapters sadvio_b200/host/b200_optimizer.hpp and sadvio_b200/api.py do this):
 *   localMapVIOptimization / localMapBA   n_fixed = fixed_frame_number, everything above zero
 *   landmarkOptimization(frame)           frames = the keyframes that see the frame's landmarks, n_fixed = n_frames (all poses
 *                                         constant, AngularAdjustmentCERESAnalytic.cpp:141-145), vio = 0, Huber sqrt(1.345),
 *                                         10 iterations; the reduced system is empty, every landmark is a 3x3 problem under
 *                                         ONE trust region
 *   singleFrameOptimization(frame)        frames = { frame }, landmarks_constant = 1, obs_sigma = 1 / focal (…Analytic.cpp:48),
 *                                         5 iterations
 *   singleFrameVIOptimization(frame)      frames = { frame, frame->getIMU()->getLastKF() }, vio = 1, one IMU pair,
 *                                         landmarks_constant = 1, Huber sqrt(1.345), 5 iterations (the 5 ms wall-clock budget
 *                                         of AOptimizer.cpp:254 is not modelled: a solve of this size takes well under 1 ms)
 * VIInit (AOptimizer.cpp:448-581) has parameter blocks of its own (gravity direction, scale): sdv_viinit below.
 */

/*
 * Solution = the values of the Ceres parameter blocks after the solve
 * (all start at zero, parametersBlock.hpp:44,84).  The caller applies them as
 * AOptimizer.cpp:391-434 does: T_f_w <- T_f_w * [exp(dw)|dt], lmk += dp,
 * v += dv, ba += dba, bg += dbg, then IMU::biasDeltaCorrection.
 */
typedef struct sdv_delta {
    double *dpose; /* [F][6]  (rotvec3, trans3) */
    double *dv;    /* [F][3]  may be NULL when !vio */
    double *dba;   /* [F][3] */
    double *dbg;   /* [F][3] */
    double *dlmk;  /* [L][3] */
} sdv_delta;

typedef struct sdv_stats {
    int32_t iterations;         /* ceres Summary::iterations.size()-1 : trust-region steps attempted */
    int32_t termination;        /* sdv_termination */
    int32_t num_successful_steps;
    int32_t num_unsuccessful_steps;
    int32_t n_reduced;          /* dimension of the dense reduced system */
    int32_t n_residual_blocks;  /* blocks kept in the reduced program */
    double initial_cost;        /* without fixed_cost, as Ceres' x_cost_ */
    double final_cost;
    double fixed_cost;          /* cost of residual blocks whose parameter blocks are all constant */
    double final_radius;
    /* per-iteration trace, entry 0 = iteration 0 */
    double trace_cost[SDV_MAX_TRACE];
    double trace_radius[SDV_MAX_TRACE];
    double trace_model_change[SDV_MAX_TRACE];
    int32_t trace_accepted[SDV_MAX_TRACE]; /* 1 accepted, 0 rejected, -1 invalid step */
    /* timing (ms): device = CUDA events around the solve (inputs resident); */
    double ms_h2d, ms_solve_device, ms_d2h, ms_total_host;
    int64_t kernel_launches;    /* kernels this call launched */
    int64_t h2d_bytes, d2h_bytes;
} sdv_stats;

typedef struct sdv_handle sdv_handle;

/* Fill `cfg` with the reference's options (AOptimizer.cpp:376-388 + Ceres 2.2 defaults). */
void sdv_default_config(sdv_config *cfg);

/* Replaces: constructing the optimizer object (slamParameters.cpp:263-282). Fails with
   SDV_ERR_NO_DEVICE when no sm_100 GPU is visible — the product has no CPU path. */
int sdv_create(sdv_handle **out, const sdv_config *cfg);
int sdv_destroy(sdv_handle *h);

/* Replaces: AOptimizer::localMapVIOptimization / localMapBA up to (not including) the
   state write-back — problem build (addResidualsLocalMap, addIMUResiduals,
   addMarginalizationResiduals) + ceres::Solve.  Host buffers in, host buffers out. */
int sdv_solve_window(sdv_handle *h, const sdv_window *win, sdv_delta *out, sdv_stats *stats);

/* The same solve split so that benchmarks can keep inputs resident in HBM:
   upload once, solve many times (each solve restarts from dx = 0), download. */
int sdv_upload_window(sdv_handle *h, const sdv_window *win);
int sdv_solve_resident(sdv_handle *h, sdv_stats *stats);
int sdv_download_delta(sdv_handle *h, sdv_delta *out);

/* Replaces: one Evaluate() sweep over every visual residual block as Ceres performs it
   (AngularAdjustmentCERESAnalytic.h:55-111 / BundleAdjustmentCERESAnalytic.h:52-90) at the
   parameter values `x` (NULL = zeros).  Outputs are per observation, Ceres row-major block
   layout: r[O][2], J_pose[O][12] (2x6), J_lmk[O][6] (2x3).  Test / profiling entry point. */
int sdv_eval_visual(sdv_handle *h, const sdv_delta *x, double *r, double *J_pose, double *J_lmk, double *cost);

/* Replaces: IMUFactor::Evaluate + IMUBiasFactor::Evaluate (residuals.hpp:133-300) for every
   pair of the uploaded window.  r_imu[P][9], J_imu[P][9*24] (row-major 9 x (6,6,3,3,3,3)
   concatenated column-wise), r_bias[P][6]. */
int sdv_eval_imu(sdv_handle *h, const sdv_delta *x, double *r_imu, double *J_imu, double *r_bias);

/*
 * IMU pre-integration — replaces: isae::IMU::processIMU (cpp/src/data/sensors/IMU.cpp:5-91) chained over
 * the IMU samples of every keyframe interval of a window: the producer of the imu_* arrays of sdv_window.
 * Sample s of an interval is the measurement held by `_last_IMU` at that step (IMU.cpp:28-31); the first sample of an
 * interval belongs to the keyframe itself (pre-integration restarts there, IMU.cpp:50-61).
 */
typedef struct sdv_imu_intervals {
    int32_t n_intervals;
    int32_t n_samples;
    const int32_t *sample_ptr; /* [n_intervals + 1]: samples of interval k are sample_ptr[k] .. sample_ptr[k+1]-1 */
    const double *acc;         /* [S][3] IMU::getAcc() of the previous measurement */
    const double *gyr;         /* [S][3] IMU::getGyr() */
    const double *dt;          /* [S]    (ts_cur - ts_last) * 1e-9; > 1 s is replaced by 1 / rate_hz (IMU.cpp:23-25) */
    const double *T_f_w;       /* [n][12] keyframe pose at the start of the interval */
    const double *v;           /* [n][3]  keyframe velocity */
    const double *ba;          /* [n][3]  keyframe biases: every measurement of the interval carries them (IMU.cpp:17-18) */
    const double *bg;          /* [n][3] */
    const double *dR_stale;    /* [n][9] or NULL (identity): the keyframe IMU's OWN _delta_R, which the reference reads for
                                          the noise matrix of the first step before restarting (IMU.cpp:44-47) */
    double eta[6];             /* IMU::_eta: (gyr_noise^2 x3, acc_noise^2 x3) * rate_hz (IMU.h:39-41) */
    double rate_hz;
} sdv_imu_intervals;

typedef struct sdv_preint { /* per interval, the state of the LAST measurement's isae::IMU; any pointer but the first five may be NULL */
    double *dR;      /* [n][9]  getDeltaR() */
    double *dv;      /* [n][3]  getDeltaV() */
    double *dp;      /* [n][3]  getDeltaP() */
    double *cov;     /* [n][81] getCov() */
    double *J_dR_bg; /* [n][9] */
    double *J_dv_ba, *J_dv_bg, *J_dp_ba, *J_dp_bg; /* [n][9] each */
    double *T_pred;  /* [n][12] dead-reckoned Frame::getWorld2FrameTransform() (IMU.cpp:38-41) */
    double *v_pred;  /* [n][3]  dead-reckoned velocity (IMU.cpp:35) */
} sdv_preint;

int sdv_preintegrate(sdv_handle *h, const sdv_imu_intervals *in, sdv_preint *out);

/*
 * Visual-inertial initialisation — replaces: the problem build + ceres::Solve of AOptimizer::VIInit
 * (cpp/src/optimizers/AOptimizer.cpp:448-529) with its functor IMUFactorInit (cpp/include/isaeslam/optimizers/residuals.hpp:302-410).
 * Read from `win`: n_frames, T_f_w, v (every frame of the local map, newest -> oldest; all of them carry an IMU) and the imu_*
 * arrays — one entry per frame j whose getIMU()->getLastKF() is another frame i of the map (AOptimizer.cpp:485-502: NO dt test
 * here); imu_J_*, imu_sigma_* are not read (the shared dba / dbg blocks are constant at zero, :472-477, and the two
 * Landmark3DPrior blocks on them, :504-515, have zero residual).  win->max_num_iterations > 0 overrides the 50 steps of :449.
 * Parameter blocks: r_wi (2), one dv per frame, lambda (constant unless optim_scale).  The caller applies the result as
 * AOptimizer.cpp:531-567 does (the adapters do): v += dv; R_w_i = Exp(r_wi0, r_wi1, 0); T_f_w <- [R_f_w R_w_i | exp(lambda) t_f_w];
 * priors re-set on the new poses; landmarks <- exp(lambda) R_w_i^T t_w_lmk.
 */
typedef struct sdv_viinit_result {
    double *dv;        /* [F][3] velocity blocks (caller-allocated) */
    double r_wi[2];    /* r_wi_par */
    double lambda;     /* log-scale; 0 when !optim_scale */
    double R_w_i[9];   /* geometry::exp_so3(r_wi[0], r_wi[1], 0), row-major */
    double scale;      /* exp(lambda): the value VIInit returns (AOptimizer.cpp:580) */
} sdv_viinit_result;

int sdv_viinit(sdv_handle *h, const sdv_window *win, int32_t optim_scale, sdv_viinit_result *out, sdv_stats *stats);

/*
 * Marginal-prior construction — replaces: AngularAdjustmentCERESAnalytic::marginalize
 * (cpp/src/optimizers/AngularAdjustmentCERESAnalytic.cpp:488-739) / BundleAdjustmentCERESAnalytic::marginalize
 * (BundleAdjustmentCERESAnalytic.cpp:431-660) with everything they call in isae::Marginalization: preMarginalize
 * (marginalization.cpp:23-143), computeInformationAndGradient (:145-211), computeSchurComplement (:213-265),
 * rankReveallingDecomposition (:318-342), computeJacobiansAndResiduals (:516-530), sparsifyVIO (:362-411), sparsifyVO (:410-514).
 *
 * `win` is the window BEFORE its oldest keyframe leaves: frame 0 of the reference = the LAST frame of `win` (frames are ordered
 * newest -> oldest), frame 1 = the one before it; `win->dense_prior` is the previous marginalisation (_marginalization_last,
 * always propagated in its dense form, …Analytic.cpp:631-660) or NULL; `win->n_fixed` is ignored (a marginalisation block has
 * no constant parameter).  The prior is over frame 1's (pose6, v3, ba3, bg3) at column 0 — VIO only — followed by the kept
 * landmarks, 3 columns each, in `keep_lmk` order (landmark indices of `win`).  Two calls: sdv_marginalize computes and reports
 * the sizes, sdv_marginal_fetch copies the result into caller-allocated buffers of those sizes.
 */
typedef struct sdv_marginal_sizes {
    int32_t ok;            /* 0: computeSchurComplement returned false (n < 4, marginalization.cpp:215) — marginalize() returns
                              false and the caller clears its marginalisation scheme (…Analytic.cpp:690-695); nothing to fetch */
    int32_t m, n;          /* _m, _n: marginalised / kept parameters */
    int32_t n_full;        /* _n_full: rows of J (eigenvalues of Ak above 1e-12) */
    int32_t n_marg, n_keep;/* landmarks marginalised / kept */
    int32_t frame;         /* index in `win` of _frame_to_keep (frame 1), -1 in the VO case */
    int32_t n_chain;       /* sparsifyVO: landmarks on the chain (n_chain - 1 LandmarkToLandmark factors) */
    int32_t eig_sweeps_m, eig_sweeps_n; /* block-Jacobi sweeps of the two eigen-decompositions */
    double ms_device;      /* CUDA events around assembly + decompositions + sparsification */
    double ms_total_host;
} sdv_marginal_sizes;

typedef struct sdv_marginal { /* caller-allocated from sdv_marginal_sizes; any pointer may be NULL (not wanted) */
    double *J;             /* [n_full][n]   _marginalization_jacobian = Lambda^1/2 U^T */
    double *r0;            /* [n_full]      _marginalization_residual = -Lambda^-1/2 U^T bk */
    int32_t *keep_lmk;     /* [n_keep]      _lmk_to_keep; column of landmark k = (frame >= 0 ? 15 : 0) + 3 k */
    int32_t *marg_lmk;     /* [n_marg]      _lmk_to_marg */
    double *Ak, *bk;       /* [n][n], [n]   _Ak, _bk */
    double *U, *Lambda;    /* [n][n_full], [n_full] */
    double *A, *b;         /* [m+n][m+n], [m+n]: information and gradient before the Schur complement (not for sdv_schur_prior) */
    /* sparsifyVIO: IMUPriordx on frame 1 (prior values = its current state) + one PoseToLandmarkFactor per kept landmark */
    double *imu_sqrt_inf;  /* [225]         _map_frame_inf */
    double *p2l_delta;     /* [n_keep][3]   _map_lmk_prior (t_f_lmk) */
    double *p2l_sqrt_inf;  /* [n_keep][9]   _map_lmk_inf */
    /* sparsifyVO: Landmark3DPrior on lmk_with_prior (prior value = its current position) + a chain of LandmarkToLandmark factors */
    int32_t *chain;        /* [n_chain]     _lmk_to_keep re-ordered by the greedy coupling walk */
    int32_t lmk_with_prior;/* out: _lmk_with_prior (landmark index of `win`), -1 if none */
    int32_t reserved0;
    double lmk_sqrt_inf[9];/* out: _info_lmk */
    double *l2l_delta;     /* [n_chain-1][3] t_k - t_k+1 */
    double *l2l_sqrt_inf;  /* [n_chain-1][9] */
} sdv_marginal;

int sdv_marginalize(sdv_handle *h, const sdv_window *win, int32_t sparsify, sdv_marginal_sizes *sizes);
int sdv_marginal_fetch(sdv_handle *h, sdv_marginal *out);
/* The dense core alone (computeSchurComplement + rankReveallingDecomposition + computeJacobiansAndResiduals) on a caller-provided
   information matrix A (row-major (m+n)^2, the m marginalised parameters first) and gradient b: the entry point the reference's
   own KAT (cpp/tests/marginalization_test.cpp:219-223, :300-313) is pinned through. */
int sdv_schur_prior(sdv_handle *h, const double *A, const double *b, int32_t m, int32_t n, double eps, sdv_marginal_sizes *sizes);

/* Multi-GPU (landmark-sharded Schur reduction, one NCCL all-reduce of [S|g|…] per LM iteration).
   `nccl_unique_id` is the 128-byte ncclUniqueId every rank received from rank 0. */
/* Host-only helper (no CUDA): the contiguous, observation-balanced landmark range [l0,l1) and observation range [o0,o1)
   that rank `rank` of `world` owns — the partition sdv_upload_window applies after sdv_comm_init. */
int sdv_shard_range(const int32_t *obs_lmk, int32_t n_obs, int32_t n_lmks, int32_t rank, int32_t world, int32_t *l0, int32_t *l1,
                    int32_t *o0, int32_t *o1);
int sdv_comm_unique_id(void *out_128_bytes);
int sdv_comm_init(sdv_handle *h, const void *nccl_unique_id, int32_t rank, int32_t world);
/* Optional, after sdv_comm_init on one NVLink / NVSwitch node (2..8 ranks): the per-iteration exchanges run as ONE kernel over
   peer memory instead of pack -> ncclAllReduce -> unpack (sadvio_b200/csrc/sdv_peer.cuh), and the whole solve stays one CUDA
   graph at N > 1.  Every rank exports its exchange area as a 64-byte cudaIpcMemHandle; the host all-gathers the handles (rank
   order) and every rank opens its peers'.  Payloads above the area's capacity keep using NCCL. */
int sdv_comm_peer_handle(sdv_handle *h, void *out_64_bytes);
int sdv_comm_peer_open(sdv_handle *h, const void *handles /* [world][64] */);

/* Benchmark helper: time `repeats` launches of one kernel of the path on the resident window with CUDA
   events on the handle's stream; returns mean ms per launch.  which: 0 materialising residual+Jacobian
   kernel (k_lin_visual), 1 fused linearisation + landmark Schur (k_lin_schur), 2 reduced-system
   factorisation + solves, 3 fused back-substitution + candidate cost; +10: cold L2 (256 MiB write between
   launches, alternating output buffers). */
int sdv_time_kernel(sdv_handle *h, int32_t which, int32_t repeats, double *ms_per_launch);

/* Test / debugging aids: copy an internal device buffer to the host (what: 0 reduced system [S|g|diag|grad],
   1 Cholesky factor, 2 reduced step, 3 jacobi scale, 4 LM damping) and report the reduced dimensions. */
int sdv_debug_read(sdv_handle *h, int32_t what, double *out, int64_t count);
int sdv_debug_dims(sdv_handle *h, int32_t *n, int32_t *n_pad);
int sdv_debug_graph_builds(sdv_handle *h, int64_t *count); /* CUDA-graph captures of this handle so far */

const char *sdv_strerror(int status);
const char *sdv_last_error(const sdv_handle *h); /* detail of the last failure on this handle */
int sdv_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SDV_H */

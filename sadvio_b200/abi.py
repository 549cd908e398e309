"""ctypes mirror of ``include/sdv.h`` plus a small numpy container for a flattened window.

Pure host-side plumbing: no compute happens here.  Field order and types MUST match include/sdv.h
(tests/test_abi.py checks sizeof/offsets against a C probe compiled from the header).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

SDV_ABI_VERSION = 2
SDV_MAX_TRACE = 64

SDV_FACTOR_ANGULAR = 0
SDV_FACTOR_PIXEL = 1

TERMINATION = {
    0: "NO_CONVERGENCE",
    1: "FUNCTION_TOLERANCE",
    2: "GRADIENT_TOLERANCE",
    3: "PARAMETER_TOLERANCE",
    4: "MIN_RADIUS",
    5: "FAILURE",
}

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)


class SdvConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("device", C.c_int32),
        ("max_num_iterations", C.c_int32),
        ("max_consecutive_invalid_steps", C.c_int32),
        ("jacobi_scaling", C.c_int32),
        ("reserved0", C.c_int32),
        ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double),
        ("parameter_tolerance", C.c_double),
        ("initial_trust_region_radius", C.c_double),
        ("max_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double),
        ("min_lm_diagonal", C.c_double),
        ("max_lm_diagonal", C.c_double),
        ("min_relative_decrease", C.c_double),
    ]


class SdvDensePrior(C.Structure):
    _fields_ = [
        ("n_full", C.c_int32),
        ("n", C.c_int32),
        ("J", c_double_p),
        ("r0", c_double_p),
        ("frame", C.c_int32),
        ("frame_col", C.c_int32),
        ("n_keep", C.c_int32),
        ("reserved0", C.c_int32),
        ("keep_lmk", c_int32_p),
        ("keep_col", c_int32_p),
    ]


class SdvSparsePrior(C.Structure):
    _fields_ = [
        ("has_imu_prior", C.c_int32),
        ("frame", C.c_int32),
        ("T_prior", C.c_double * 12),
        ("v_prior", C.c_double * 3),
        ("ba_prior", C.c_double * 3),
        ("bg_prior", C.c_double * 3),
        ("imu_sqrt_inf", C.c_double * 225),
        ("n_p2l", C.c_int32),
        ("reserved0", C.c_int32),
        ("p2l_lmk", c_int32_p),
        ("p2l_delta", c_double_p),
        ("p2l_sqrt_inf", c_double_p),
        ("has_lmk_prior", C.c_int32),
        ("lmk0", C.c_int32),
        ("lmk_prior", C.c_double * 3),
        ("lmk_sqrt_inf", C.c_double * 9),
        ("n_l2l", C.c_int32),
        ("reserved1", C.c_int32),
        ("l2l_a", c_int32_p),
        ("l2l_b", c_int32_p),
        ("l2l_delta", c_double_p),
        ("l2l_sqrt_inf", c_double_p),
    ]


class SdvWindow(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("vio", C.c_int32),
        ("factor_kind", C.c_int32),
        ("n_frames", C.c_int32),
        ("n_fixed", C.c_int32),
        ("n_cams", C.c_int32),
        ("n_lmks", C.c_int32),
        ("n_obs", C.c_int32),
        ("n_imu", C.c_int32),
        ("reserved0", C.c_int32),
        ("T_f_w", c_double_p),
        ("v", c_double_p),
        ("ba", c_double_p),
        ("bg", c_double_p),
        ("has_imu", c_uint8_p),
        ("has_prior", c_uint8_p),
        ("T_prior", c_double_p),
        ("inf_prior", c_double_p),
        ("T_s_f", c_double_p),
        ("K", c_double_p),
        ("lmk_t", c_double_p),
        ("obs_lmk", c_int32_p),
        ("obs_frame", c_int32_p),
        ("obs_cam", c_int32_p),
        ("obs_bearing", c_double_p),
        ("obs_uv", c_double_p),
        ("obs_sigma", c_double_p),
        ("imu_i", c_int32_p),
        ("imu_j", c_int32_p),
        ("imu_dt", c_double_p),
        ("imu_dR", c_double_p),
        ("imu_dv", c_double_p),
        ("imu_dp", c_double_p),
        ("imu_cov", c_double_p),
        ("imu_J_dR_bg", c_double_p),
        ("imu_J_dv_ba", c_double_p),
        ("imu_J_dv_bg", c_double_p),
        ("imu_J_dp_ba", c_double_p),
        ("imu_J_dp_bg", c_double_p),
        ("imu_sigma_ba", c_double_p),
        ("imu_sigma_bg", c_double_p),
        ("dense_prior", C.POINTER(SdvDensePrior)),
        ("sparse_prior", C.POINTER(SdvSparsePrior)),
        ("visual_loss_huber_a", C.c_double),
        ("landmarks_constant", C.c_int32),
        ("max_num_iterations", C.c_int32),
        ("lmk_has_prior", C.POINTER(C.c_uint8)),
    ]


class SdvDelta(C.Structure):
    _fields_ = [
        ("dpose", c_double_p),
        ("dv", c_double_p),
        ("dba", c_double_p),
        ("dbg", c_double_p),
        ("dlmk", c_double_p),
    ]


class SdvStats(C.Structure):
    _fields_ = [
        ("iterations", C.c_int32),
        ("termination", C.c_int32),
        ("num_successful_steps", C.c_int32),
        ("num_unsuccessful_steps", C.c_int32),
        ("n_reduced", C.c_int32),
        ("n_residual_blocks", C.c_int32),
        ("initial_cost", C.c_double),
        ("final_cost", C.c_double),
        ("fixed_cost", C.c_double),
        ("final_radius", C.c_double),
        ("trace_cost", C.c_double * SDV_MAX_TRACE),
        ("trace_radius", C.c_double * SDV_MAX_TRACE),
        ("trace_model_change", C.c_double * SDV_MAX_TRACE),
        ("trace_accepted", C.c_int32 * SDV_MAX_TRACE),
        ("ms_h2d", C.c_double),
        ("ms_solve_device", C.c_double),
        ("ms_d2h", C.c_double),
        ("ms_total_host", C.c_double),
        ("kernel_launches", C.c_int64),
        ("h2d_bytes", C.c_int64),
        ("d2h_bytes", C.c_int64),
    ]


class SdvMarginalSizes(C.Structure):
    _fields_ = [
        ("ok", C.c_int32), ("m", C.c_int32), ("n", C.c_int32), ("n_full", C.c_int32), ("n_marg", C.c_int32), ("n_keep", C.c_int32),
        ("frame", C.c_int32), ("n_chain", C.c_int32), ("eig_sweeps_m", C.c_int32), ("eig_sweeps_n", C.c_int32),
        ("ms_device", C.c_double), ("ms_total_host", C.c_double),
    ]


class SdvMarginal(C.Structure):
    _fields_ = [
        ("J", c_double_p), ("r0", c_double_p), ("keep_lmk", c_int32_p), ("marg_lmk", c_int32_p), ("Ak", c_double_p), ("bk", c_double_p),
        ("U", c_double_p), ("Lambda", c_double_p), ("A", c_double_p), ("b", c_double_p),
        ("imu_sqrt_inf", c_double_p), ("p2l_delta", c_double_p), ("p2l_sqrt_inf", c_double_p),
        ("chain", c_int32_p), ("lmk_with_prior", C.c_int32), ("reserved0", C.c_int32), ("lmk_sqrt_inf", C.c_double * 9),
        ("l2l_delta", c_double_p), ("l2l_sqrt_inf", c_double_p),
    ]


class SdvImuIntervals(C.Structure):
    _fields_ = [
        ("n_intervals", C.c_int32),
        ("n_samples", C.c_int32),
        ("sample_ptr", c_int32_p),
        ("acc", c_double_p),
        ("gyr", c_double_p),
        ("dt", c_double_p),
        ("T_f_w", c_double_p),
        ("v", c_double_p),
        ("ba", c_double_p),
        ("bg", c_double_p),
        ("dR_stale", c_double_p),
        ("eta", C.c_double * 6),
        ("rate_hz", C.c_double),
    ]


class SdvPreint(C.Structure):
    _fields_ = [(n, c_double_p) for n in ("dR", "dv", "dp", "cov", "J_dR_bg", "J_dv_ba", "J_dv_bg", "J_dp_ba", "J_dp_bg", "T_pred", "v_pred")]


class SdvViinitResult(C.Structure):
    _fields_ = [("dv", c_double_p), ("r_wi", C.c_double * 2), ("lambda_", C.c_double), ("R_w_i", C.c_double * 9), ("scale", C.c_double)]


def _dp(a: Optional[np.ndarray]):
    if a is None:
        return c_double_p()
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(c_double_p)


def _ip(a: Optional[np.ndarray]):
    if a is None:
        return c_int32_p()
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_int32_p)


def _up(a: Optional[np.ndarray]):
    if a is None:
        return c_uint8_p()
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_uint8_p)


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


@dataclass
class DensePrior:
    """isae::MarginalizationFactor inputs (marginalization.hpp:88-218)."""

    J: np.ndarray          # [n_full, n]
    r0: np.ndarray         # [n_full]
    frame: int             # index of frame_to_keep or -1
    frame_col: int
    keep_lmk: np.ndarray   # [n_keep] int32
    keep_col: np.ndarray   # [n_keep] int32


@dataclass
class SparsePrior:
    """Sparsified prior inputs (AngularAdjustmentCERESAnalytic.cpp:387-483)."""

    has_imu_prior: bool = False
    frame: int = -1
    T_prior: Optional[np.ndarray] = None
    v_prior: Optional[np.ndarray] = None
    ba_prior: Optional[np.ndarray] = None
    bg_prior: Optional[np.ndarray] = None
    imu_sqrt_inf: Optional[np.ndarray] = None
    p2l_lmk: Optional[np.ndarray] = None
    p2l_delta: Optional[np.ndarray] = None
    p2l_sqrt_inf: Optional[np.ndarray] = None
    has_lmk_prior: bool = False
    lmk0: int = -1
    lmk_prior: Optional[np.ndarray] = None
    lmk_sqrt_inf: Optional[np.ndarray] = None
    l2l_a: Optional[np.ndarray] = None
    l2l_b: Optional[np.ndarray] = None
    l2l_delta: Optional[np.ndarray] = None
    l2l_sqrt_inf: Optional[np.ndarray] = None


@dataclass
class SkippedPreint:
    """Pre-integration of the frames that get NO IMUFactor although their getLastKF() is in the window with an IMU
    (dt > 1 s, AOptimizer.cpp:69, or framei == framej, :72).  They never cross the C ABI — the solve does not see them —
    but the write-back still corrects their deltas with the previous keyframe's dba / dbg: the loop at
    AOptimizer.cpp:421-434 has neither test."""

    frame: np.ndarray      # [K] window index of the frame
    prev: np.ndarray       # [K] window index of its getLastKF()
    dR: np.ndarray         # [K,9]
    dv: np.ndarray         # [K,3]
    dp: np.ndarray         # [K,3]
    J_dR_bg: np.ndarray    # [K,9]
    J_dv_ba: np.ndarray
    J_dv_bg: np.ndarray
    J_dp_ba: np.ndarray
    J_dp_bg: np.ndarray


@dataclass
class Window:
    """A flattened sliding window (numpy SoA). ``as_struct`` yields the C view (keeps arrays alive)."""

    vio: bool
    factor_kind: int
    n_fixed: int
    T_f_w: np.ndarray                  # [F,12]
    T_s_f: np.ndarray                  # [C,12]
    K: np.ndarray                      # [C,4]
    lmk_t: np.ndarray                  # [L,3]
    obs_lmk: np.ndarray                # [O]
    obs_frame: np.ndarray
    obs_cam: np.ndarray
    obs_bearing: Optional[np.ndarray] = None   # [O,3]
    obs_uv: Optional[np.ndarray] = None        # [O,2]
    obs_sigma: Optional[np.ndarray] = None
    v: Optional[np.ndarray] = None
    ba: Optional[np.ndarray] = None
    bg: Optional[np.ndarray] = None
    has_imu: Optional[np.ndarray] = None
    has_prior: Optional[np.ndarray] = None
    T_prior: Optional[np.ndarray] = None
    inf_prior: Optional[np.ndarray] = None
    imu_i: Optional[np.ndarray] = None
    imu_j: Optional[np.ndarray] = None
    imu_dt: Optional[np.ndarray] = None
    imu_dR: Optional[np.ndarray] = None
    imu_dv: Optional[np.ndarray] = None
    imu_dp: Optional[np.ndarray] = None
    imu_cov: Optional[np.ndarray] = None
    imu_J_dR_bg: Optional[np.ndarray] = None
    imu_J_dv_ba: Optional[np.ndarray] = None
    imu_J_dv_bg: Optional[np.ndarray] = None
    imu_J_dp_ba: Optional[np.ndarray] = None
    imu_J_dp_bg: Optional[np.ndarray] = None
    imu_sigma_ba: Optional[np.ndarray] = None
    imu_sigma_bg: Optional[np.ndarray] = None
    dense_prior: Optional[DensePrior] = None
    sparse_prior: Optional[SparsePrior] = None
    skipped_preint: Optional[SkippedPreint] = None   # host-side only (write-back), see SkippedPreint
    visual_loss_huber_a: float = 0.0           # ceres::HuberLoss(a) on the visual residual blocks, 0 = none
    landmarks_constant: bool = False           # single-frame solves: every landmark block constant
    max_num_iterations: int = 0                # > 0: overrides the configuration for this window
    lmk_has_prior: Optional[np.ndarray] = None # [L] uint8, ALandmark::hasPrior() (sdv_marginalize only)
    meta: dict = field(default_factory=dict)   # ground truth etc. (never crosses the ABI)

    @property
    def n_frames(self) -> int:
        return int(self.T_f_w.shape[0])

    @property
    def n_lmks(self) -> int:
        return int(self.lmk_t.shape[0])

    @property
    def n_obs(self) -> int:
        return int(self.obs_lmk.shape[0])

    @property
    def n_imu(self) -> int:
        return 0 if self.imu_i is None else int(self.imu_i.shape[0])

    def normalise(self) -> "Window":
        F = self.n_frames
        self.T_f_w = _f64(self.T_f_w, (F, 12))
        self.T_s_f = _f64(self.T_s_f, (-1, 12))
        self.K = _f64(self.K, (-1, 4))
        self.lmk_t = _f64(self.lmk_t, (-1, 3))
        for name in ("obs_lmk", "obs_frame", "obs_cam", "imu_i", "imu_j"):
            setattr(self, name, _i32(getattr(self, name)))
        for name in (
            "obs_bearing", "obs_uv", "obs_sigma", "v", "ba", "bg", "T_prior", "inf_prior", "imu_dt", "imu_dR",
            "imu_dv", "imu_dp", "imu_cov", "imu_J_dR_bg", "imu_J_dv_ba", "imu_J_dv_bg", "imu_J_dp_ba",
            "imu_J_dp_bg", "imu_sigma_ba", "imu_sigma_bg",
        ):
            setattr(self, name, _f64(getattr(self, name)))
        for name in ("has_imu", "has_prior"):
            a = getattr(self, name)
            if a is not None:
                setattr(self, name, np.ascontiguousarray(a, dtype=np.uint8))
        return self

    def as_struct(self) -> SdvWindow:
        self.normalise()
        w = SdvWindow()
        w.abi_version = SDV_ABI_VERSION
        w.vio = 1 if self.vio else 0
        w.factor_kind = int(self.factor_kind)
        w.n_frames = self.n_frames
        w.n_fixed = int(self.n_fixed)
        w.n_cams = int(self.T_s_f.shape[0])
        w.n_lmks = self.n_lmks
        w.n_obs = self.n_obs
        w.n_imu = self.n_imu
        w.T_f_w = _dp(self.T_f_w)
        w.v, w.ba, w.bg = _dp(self.v), _dp(self.ba), _dp(self.bg)
        w.has_imu = _up(self.has_imu)
        w.has_prior = _up(self.has_prior)
        w.T_prior = _dp(self.T_prior)
        w.inf_prior = _dp(self.inf_prior)
        w.T_s_f = _dp(self.T_s_f)
        w.K = _dp(self.K)
        w.lmk_t = _dp(self.lmk_t)
        w.obs_lmk, w.obs_frame, w.obs_cam = _ip(self.obs_lmk), _ip(self.obs_frame), _ip(self.obs_cam)
        w.obs_bearing = _dp(self.obs_bearing)
        w.obs_uv = _dp(self.obs_uv)
        w.obs_sigma = _dp(self.obs_sigma)
        w.imu_i, w.imu_j = _ip(self.imu_i), _ip(self.imu_j)
        for name in (
            "imu_dt", "imu_dR", "imu_dv", "imu_dp", "imu_cov", "imu_J_dR_bg", "imu_J_dv_ba", "imu_J_dv_bg",
            "imu_J_dp_ba", "imu_J_dp_bg", "imu_sigma_ba", "imu_sigma_bg",
        ):
            setattr(w, name, _dp(getattr(self, name)))
        w.visual_loss_huber_a = float(self.visual_loss_huber_a)
        w.landmarks_constant = 1 if self.landmarks_constant else 0
        w.max_num_iterations = int(self.max_num_iterations)
        if self.lmk_has_prior is not None:
            self.lmk_has_prior = np.ascontiguousarray(self.lmk_has_prior, dtype=np.uint8)
            w.lmk_has_prior = self.lmk_has_prior.ctypes.data_as(C.POINTER(C.c_uint8))
        keep = [self]
        if self.dense_prior is not None:
            d = self.dense_prior
            d.J = _f64(d.J)
            d.r0 = _f64(d.r0)
            d.keep_lmk = _i32(d.keep_lmk)
            d.keep_col = _i32(d.keep_col)
            s = SdvDensePrior()
            s.n_full, s.n = int(d.J.shape[0]), int(d.J.shape[1])
            s.J, s.r0 = _dp(d.J), _dp(d.r0)
            s.frame, s.frame_col = int(d.frame), int(d.frame_col)
            s.n_keep = int(d.keep_lmk.shape[0])
            s.keep_lmk, s.keep_col = _ip(d.keep_lmk), _ip(d.keep_col)
            keep.append(s)
            w.dense_prior = C.pointer(s)
        if self.sparse_prior is not None:
            p = self.sparse_prior
            s = SdvSparsePrior()
            s.has_imu_prior = 1 if p.has_imu_prior else 0
            s.frame = int(p.frame)
            if p.has_imu_prior:
                s.T_prior[:] = list(np.asarray(p.T_prior, dtype=np.float64).reshape(12))
                s.v_prior[:] = list(np.asarray(p.v_prior, dtype=np.float64).reshape(3))
                s.ba_prior[:] = list(np.asarray(p.ba_prior, dtype=np.float64).reshape(3))
                s.bg_prior[:] = list(np.asarray(p.bg_prior, dtype=np.float64).reshape(3))
                s.imu_sqrt_inf[:] = list(np.asarray(p.imu_sqrt_inf, dtype=np.float64).reshape(225))
            p.p2l_lmk = _i32(p.p2l_lmk)
            p.p2l_delta = _f64(p.p2l_delta)
            p.p2l_sqrt_inf = _f64(p.p2l_sqrt_inf)
            s.n_p2l = 0 if p.p2l_lmk is None else int(p.p2l_lmk.shape[0])
            s.p2l_lmk, s.p2l_delta, s.p2l_sqrt_inf = _ip(p.p2l_lmk), _dp(p.p2l_delta), _dp(p.p2l_sqrt_inf)
            s.has_lmk_prior = 1 if p.has_lmk_prior else 0
            s.lmk0 = int(p.lmk0)
            if p.has_lmk_prior:
                s.lmk_prior[:] = list(np.asarray(p.lmk_prior, dtype=np.float64).reshape(3))
                s.lmk_sqrt_inf[:] = list(np.asarray(p.lmk_sqrt_inf, dtype=np.float64).reshape(9))
            p.l2l_a, p.l2l_b = _i32(p.l2l_a), _i32(p.l2l_b)
            p.l2l_delta, p.l2l_sqrt_inf = _f64(p.l2l_delta), _f64(p.l2l_sqrt_inf)
            s.n_l2l = 0 if p.l2l_a is None else int(p.l2l_a.shape[0])
            s.l2l_a, s.l2l_b = _ip(p.l2l_a), _ip(p.l2l_b)
            s.l2l_delta, s.l2l_sqrt_inf = _dp(p.l2l_delta), _dp(p.l2l_sqrt_inf)
            keep.append(s)
            w.sparse_prior = C.pointer(s)
        w._keepalive = keep  # noqa: attribute on the ctypes instance keeps numpy buffers alive
        return w


@dataclass
class Delta:
    """Solution blocks (numpy), mirrors sdv_delta."""

    dpose: np.ndarray
    dv: np.ndarray
    dba: np.ndarray
    dbg: np.ndarray
    dlmk: np.ndarray

    @staticmethod
    def zeros(n_frames: int, n_lmks: int) -> "Delta":
        return Delta(
            np.zeros((n_frames, 6)), np.zeros((n_frames, 3)), np.zeros((n_frames, 3)), np.zeros((n_frames, 3)),
            np.zeros((n_lmks, 3)),
        )

    def as_struct(self) -> SdvDelta:
        d = SdvDelta()
        for name in ("dpose", "dv", "dba", "dbg", "dlmk"):
            a = np.ascontiguousarray(getattr(self, name), dtype=np.float64)
            setattr(self, name, a)
            setattr(d, name, _dp(a))
        d._keepalive = self
        return d


def stats_to_dict(st: SdvStats) -> dict:
    n = min(int(st.iterations) + 1, SDV_MAX_TRACE)
    return {
        "iterations": int(st.iterations),
        "termination": TERMINATION.get(int(st.termination), str(st.termination)),
        "num_successful_steps": int(st.num_successful_steps),
        "num_unsuccessful_steps": int(st.num_unsuccessful_steps),
        "n_reduced": int(st.n_reduced),
        "n_residual_blocks": int(st.n_residual_blocks),
        "initial_cost": float(st.initial_cost),
        "final_cost": float(st.final_cost),
        "fixed_cost": float(st.fixed_cost),
        "final_radius": float(st.final_radius),
        "trace_cost": [float(st.trace_cost[i]) for i in range(n)],
        "trace_radius": [float(st.trace_radius[i]) for i in range(n)],
        "trace_model_change": [float(st.trace_model_change[i]) for i in range(n)],
        "trace_accepted": [int(st.trace_accepted[i]) for i in range(n)],
        "ms_h2d": float(st.ms_h2d),
        "ms_solve_device": float(st.ms_solve_device),
        "ms_d2h": float(st.ms_d2h),
        "ms_total_host": float(st.ms_total_host),
        "kernel_launches": int(st.kernel_launches),
        "h2d_bytes": int(st.h2d_bytes),
        "d2h_bytes": int(st.d2h_bytes),
    }

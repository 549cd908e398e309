"""ctypes loader for the CPU oracle.  TEST INFRASTRUCTURE.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this module; nothing under ``sadvio_b200/`` does (tests/test_layout.py greps for it).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from sadvio_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_SRCS = ["sdv_oracle.cpp", "factors.hpp", "marg.hpp", "smallmat.hpp", os.path.join("..", "include", "sdv.h")]


def build(force: bool = False) -> str:
    stale = force or not os.path.exists(_SO)
    if not stale:
        t = os.path.getmtime(_SO)
        stale = any(os.path.getmtime(os.path.join(_HERE, s)) > t for s in _SRCS)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        dp, ip = abi.c_double_p, abi.c_int32_p
        _lib.orc_default_config.argtypes = [C.POINTER(abi.SdvConfig)]
        _lib.orc_solve_window.argtypes = [C.POINTER(abi.SdvWindow), C.POINTER(abi.SdvConfig), C.POINTER(abi.SdvDelta),
                                          C.POINTER(abi.SdvStats), C.c_int, C.c_int]
        _lib.orc_eval_visual.argtypes = [C.POINTER(abi.SdvWindow), C.POINTER(abi.SdvDelta), dp, dp, dp, dp]
        _lib.orc_eval_imu.argtypes = [C.POINTER(abi.SdvWindow), C.POINTER(abi.SdvDelta), dp, dp, dp]
        _lib.orc_reduced_system.argtypes = [C.POINTER(abi.SdvWindow), C.c_int, C.c_int, C.c_int, C.c_double, dp, dp, C.POINTER(C.c_int)]
        _lib.orc_cost.argtypes = [C.POINTER(abi.SdvWindow), C.POINTER(abi.SdvDelta), dp, dp]
        _lib.orc_schur_prior.argtypes = [C.c_int, C.c_int, dp, dp, C.c_double, dp, dp, C.POINTER(C.c_int), dp, dp, dp, dp]
        _lib.orc_schur_prior.restype = C.c_int
        for name in ("orc_exp_so3", "orc_log_so3", "orc_right_jacobian"):
            getattr(_lib, name).argtypes = [dp, dp]
        _lib.orc_angular_eval.argtypes = [dp, dp, dp, dp, C.c_double, dp, dp, dp, dp, dp]
        _lib.orc_reproj_eval.argtypes = [dp, dp, dp, dp, dp, C.c_double, dp, dp, dp, dp, dp]
        _lib.orc_pose_prior_eval.argtypes = [dp, dp, dp, dp, dp, dp]
        _lib.orc_p2l_eval.argtypes = [dp, dp, dp, dp, dp, dp, dp, dp, dp]
        _lib.orc_imu_factor_eval.argtypes = [dp, dp, dp, dp, C.c_double, dp, dp, dp, dp]
        _lib.orc_imu_init_eval.argtypes = [dp, dp, dp, dp, C.c_double, dp, dp, dp, dp]
        _lib.orc_viinit.argtypes = [C.POINTER(abi.SdvWindow), C.POINTER(abi.SdvConfig), C.c_int, dp, dp, C.POINTER(abi.SdvStats)]
        _lib.orc_process_imu.argtypes = [dp, dp, dp, C.c_double, dp, C.c_double, dp]
        _lib.orc_bias_delta_correction.argtypes = [dp, dp, dp]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(abi.c_double_p)


def _a(x, n=None):
    a = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
    if n is not None:
        assert a.size == n, (a.size, n)
    return a


def default_config() -> abi.SdvConfig:
    cfg = abi.SdvConfig()
    lib().orc_default_config(C.byref(cfg))
    return cfg


def solve_window(win: abi.Window, cfg: abi.SdvConfig | None = None, mode: int = 0, nthreads: int = 1):
    """Run the restated Ceres LM loop. mode 0 = landmark Schur elimination, 1 = dense full normal equations."""
    cfg = cfg or default_config()
    ws = win.as_struct()
    out = abi.Delta.zeros(win.n_frames, win.n_lmks)
    ds = out.as_struct()
    st = abi.SdvStats()
    rc = lib().orc_solve_window(C.byref(ws), C.byref(cfg), C.byref(ds), C.byref(st), mode, nthreads)
    return rc, out, abi.stats_to_dict(st)


def eval_visual(win: abi.Window, x: abi.Delta | None = None):
    ws = win.as_struct()
    O = win.n_obs
    r, Jp, Jl = np.zeros((O, 2)), np.zeros((O, 12)), np.zeros((O, 6))
    cost = np.zeros(1)
    xs = x.as_struct() if x is not None else None
    lib().orc_eval_visual(C.byref(ws), C.byref(xs) if xs is not None else None, _p(r), _p(Jp), _p(Jl), _p(cost))
    return r, Jp, Jl, float(cost[0])


def eval_imu(win: abi.Window, x: abi.Delta | None = None):
    ws = win.as_struct()
    P = win.n_imu
    r, J, rb = np.zeros((P, 9)), np.zeros((P, 9 * 24)), np.zeros((P, 6))
    xs = x.as_struct() if x is not None else None
    lib().orc_eval_imu(C.byref(ws), C.byref(xs) if xs is not None else None, _p(r), _p(J), _p(rb))
    return r, J.reshape(P, 9, 24), rb


def cost(win: abi.Window, x: abi.Delta | None = None):
    ws = win.as_struct()
    c, fc = np.zeros(1), np.zeros(1)
    xs = x.as_struct() if x is not None else None
    lib().orc_cost(C.byref(ws), C.byref(xs) if xs is not None else None, _p(c), _p(fc))
    return float(c[0]), float(fc[0])


def reduced_system(win: abi.Window, l0: int, l1: int, with_factors: bool, lam: float = 1e-4):
    """Partial reduced (Schur) system of the landmark range [l0, l1) at x = 0."""
    ws = win.as_struct()
    n = C.c_int(0)
    lib().orc_reduced_system(C.byref(ws), l0, l1, int(with_factors), lam, None, None, C.byref(n))
    S, g = np.zeros((n.value, n.value)), np.zeros(n.value)
    lib().orc_reduced_system(C.byref(ws), l0, l1, int(with_factors), lam, _p(S), _p(g), C.byref(n))
    return S, g


def exp_so3(v):
    R = np.zeros(9)
    lib().orc_exp_so3(_p(_a(v, 3)), _p(R))
    return R.reshape(3, 3)


def log_so3(R):
    v = np.zeros(3)
    lib().orc_log_so3(_p(_a(R, 9)), _p(v))
    return v


def right_jacobian(v):
    J = np.zeros(9)
    lib().orc_right_jacobian(_p(_a(v, 3)), _p(J))
    return J.reshape(3, 3)


def angular_eval(bearing, T_s_f, T_f_w, t_w_lmk, sigma, dx=None, dp=None, jac=True):
    dx = _a(np.zeros(6) if dx is None else dx, 6)
    dp = _a(np.zeros(3) if dp is None else dp, 3)
    r, J6, J3 = np.zeros(2), np.zeros(12), np.zeros(6)
    lib().orc_angular_eval(_p(_a(bearing, 3)), _p(_a(T_s_f, 12)), _p(_a(T_f_w, 12)), _p(_a(t_w_lmk, 3)), float(sigma),
                           _p(dx), _p(dp), _p(r), _p(J6) if jac else None, _p(J3) if jac else None)
    return r, J6.reshape(2, 6), J3.reshape(2, 3)


def reproj_eval(uv, K, T_s_f, T_f_w, t_w_lmk, sigma=1.0, dx=None, dp=None, jac=True):
    dx = _a(np.zeros(6) if dx is None else dx, 6)
    dp = _a(np.zeros(3) if dp is None else dp, 3)
    r, J6, J3 = np.zeros(2), np.zeros(12), np.zeros(6)
    lib().orc_reproj_eval(_p(_a(uv, 2)), _p(_a(K, 4)), _p(_a(T_s_f, 12)), _p(_a(T_f_w, 12)), _p(_a(t_w_lmk, 3)),
                          float(sigma), _p(dx), _p(dp), _p(r), _p(J6) if jac else None, _p(J3) if jac else None)
    return r, J6.reshape(2, 6), J3.reshape(2, 3)


def pose_prior_eval(T, T_prior, sqrt_inf_diag, dx=None, jac=True):
    dx = _a(np.zeros(6) if dx is None else dx, 6)
    r, J = np.zeros(6), np.zeros(36)
    lib().orc_pose_prior_eval(_p(_a(T, 12)), _p(_a(T_prior, 12)), _p(_a(sqrt_inf_diag, 6)), _p(dx), _p(r),
                              _p(J) if jac else None)
    return r, J.reshape(6, 6)


def p2l_eval(delta, T_f_w, t_w_lmk, sqrt_inf, dx=None, dp=None, jac=True):
    dx = _a(np.zeros(6) if dx is None else dx, 6)
    dp = _a(np.zeros(3) if dp is None else dp, 3)
    r, J6, J3 = np.zeros(3), np.zeros(18), np.zeros(9)
    lib().orc_p2l_eval(_p(_a(delta, 3)), _p(_a(T_f_w, 12)), _p(_a(t_w_lmk, 3)), _p(_a(sqrt_inf, 9)), _p(dx), _p(dp),
                       _p(r), _p(J6) if jac else None, _p(J3) if jac else None)
    return r, J6.reshape(3, 6), J3.reshape(3, 3)


def pack_preint(dR, dv, dp, cov, J_dR_bg, J_dv_ba, J_dv_bg, J_dp_ba, J_dp_bg):
    return np.concatenate([_a(dR, 9), _a(dv, 3), _a(dp, 3), _a(cov, 81), _a(J_dR_bg, 9), _a(J_dv_ba, 9), _a(J_dv_bg, 9),
                           _a(J_dp_ba, 9), _a(J_dp_bg, 9)])


def imu_factor_eval(T_i, T_j, v_i, v_j, dt, pre, params=None, jac=True):
    params = _a(np.zeros(24) if params is None else params, 24)
    r, J = np.zeros(9), np.zeros(9 * 24)
    rc = lib().orc_imu_factor_eval(_p(_a(T_i, 12)), _p(_a(T_j, 12)), _p(_a(v_i, 3)), _p(_a(v_j, 3)), float(dt),
                                   _p(_a(pre, 141)), _p(params), _p(r), _p(J) if jac else None)
    assert rc == 0
    return r, J.reshape(9, 24)


def imu_init_eval(T_i, T_j, v_i, v_j, dt, pre, params=None, jac=True):
    """IMUFactorInit::Evaluate (residuals.hpp:302-410); params = (w_x, w_y, dv_i, dv_j, dba, dbg, lambda), J is 9 x 15."""
    params = _a(np.zeros(15) if params is None else params, 15)
    r, J = np.zeros(9), np.zeros(9 * 15)
    rc = lib().orc_imu_init_eval(_p(_a(T_i, 12)), _p(_a(T_j, 12)), _p(_a(v_i, 3)), _p(_a(v_j, 3)), float(dt),
                                 _p(_a(pre, 141)), _p(params), _p(r), _p(J) if jac else None)
    assert rc == 0
    return r, J.reshape(9, 15)


def viinit(win: abi.Window, optim_scale: bool = True, cfg: abi.SdvConfig | None = None, all_blocks_free: bool = False):
    """The solve of AOptimizer::VIInit (AOptimizer.cpp:448-529) over the frames / IMU pairs of `win`.
    Returns rc, dict(dv[F][3], r_wi[2], lam, R_w_i, scale), stats.  all_blocks_free: the shared dba / dbg blocks are parameters too
    (the reference's own test of the functor, imu_test.cpp:498-541; VIInit itself sets them constant)."""
    cfg = cfg or default_config()
    ws = win.as_struct()
    dv, extra = np.zeros((win.n_frames, 3)), np.zeros(3)
    st = abi.SdvStats()
    rc = lib().orc_viinit(C.byref(ws), C.byref(cfg), 2 if all_blocks_free else int(bool(optim_scale)), _p(dv), _p(extra), C.byref(st))
    return rc, dict(dv=dv, r_wi=extra[:2].copy(), lam=float(extra[2]), R_w_i=exp_so3([extra[0], extra[1], 0.0]),
                    scale=float(np.exp(extra[2]))), abi.stats_to_dict(st)


IMU_STATE = dict(acc=(0, 3), gyr=(3, 6), ba=(6, 9), bg=(9, 12), v=(12, 15), T_f_w=(15, 27), is_kf=(27, 28),
                 dR=(28, 37), dv=(37, 40), dp=(40, 43), Sigma=(43, 124), J_dR_bg=(124, 133), J_dv_ba=(133, 142),
                 J_dv_bg=(142, 151), J_dp_ba=(151, 160), J_dp_bg=(160, 169))


def imu_state(acc, gyr, T_f_w=None, v=None, ba=None, bg=None, is_kf=False):
    """An isae::IMU object as the config constructor leaves it (IMU.h:26-42): zero biases/velocity, delta_R = I."""
    s = np.zeros(169)
    s[0:3], s[3:6] = acc, gyr
    s[6:9] = 0 if ba is None else ba
    s[9:12] = 0 if bg is None else bg
    s[12:15] = 0 if v is None else v
    s[15:27] = np.eye(3, 4).reshape(12) if T_f_w is None else _a(T_f_w, 12)
    s[27] = 1.0 if is_kf else 0.0
    s[28:37] = np.eye(3).reshape(9)
    return s


def imu_get(s, name):
    a, b = IMU_STATE[name]
    return s[a:b].copy()


def process_imu(last, kf_ba, kf_bg, dt, eta, rate_hz, acc, gyr, is_kf=False):
    cur = imu_state(acc, gyr, is_kf=is_kf)
    lib().orc_process_imu(_p(_a(last, 169)), _p(_a(kf_ba, 3)), _p(_a(kf_bg, 3)), float(dt), _p(_a(eta, 6)), float(rate_hz),
                          _p(cur))
    return cur


def bias_delta_correction(state, d_ba, d_bg):
    s = _a(state, 169).copy()
    lib().orc_bias_delta_correction(_p(s), _p(_a(d_ba, 3)), _p(_a(d_bg, 3)))
    return s


def sym_eig(A, method: str = "ql"):
    """Eigen-decomposition of a symmetric matrix with the oracle's own solvers (marg.hpp): "ql" = Householder + implicit QL,
    "jacobi" = cyclic Jacobi.  Returns (w ascending, V with eigenvectors in columns)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    w, V = np.zeros(n), np.zeros((n, n))
    L = lib()
    L.orc_sym_eig.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.orc_sym_eig(n, _p(A), 0 if method == "ql" else 1, _p(w), _p(V))
    return w, V


def schur_prior(A, b, m: int, eps: float = 1e-12):
    """Dense core of the reference's marginal-prior construction (marginalization.cpp:213-265, 318-342, 516-530): the first m
    parameters of the information matrix A / gradient b are marginalised.  Returns None when the reference returns false
    (fewer than 4 kept parameters), else a dict with Ak, bk, U, Lambda, J (= Lambda^1/2 U^T) and r0 (= -Lambda^-1/2 U^T bk)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    N = A.shape[0]
    n = N - m
    Ak, bk, U, Lam, J, r0 = np.zeros((max(n, 1), max(n, 1))), np.zeros(max(n, 1)), np.zeros(max(n, 1) ** 2), np.zeros(max(n, 1)), np.zeros(max(n, 1) ** 2), np.zeros(max(n, 1))
    nf = C.c_int(0)
    ok = lib().orc_schur_prior(m, n, _p(A), _p(b), eps, _p(Ak), _p(bk), C.byref(nf), _p(U), _p(Lam), _p(J), _p(r0))
    if not ok:
        return None
    k = nf.value
    return {"Ak": Ak, "bk": bk, "n_full": k, "U": U[:n * k].reshape(n, k).copy(), "Lambda": Lam[:k].copy(), "J": J[:k * n].reshape(k, n).copy(), "r0": r0[:k].copy()}

"""The other three visual solves of AOptimizer as masks of the window solve:
landmarkOptimization (AOptimizer.cpp:98-150: every pose constant, ceres::HuberLoss(sqrt(1.345)), 10 iterations),
singleFrameOptimization (:152-217: landmarks constant, sigma 1 / focal, 5 iterations) and singleFrameVIOptimization
(:219-297: the same + the IMU factor to the previous keyframe, Huber on the visual blocks only).

CPU tests pin the oracle's restatement of the loss (Ceres 2.2 HuberLoss + Corrector are not part of /root/reference) on
properties that do not depend on the oracle's own evaluation code: the reported cost equals 1/2 sum rho(|r|^2) formed in
numpy from the raw residuals, and the converged point is a stationary point of that robust cost.  GPU tests compare the CUDA
path with the oracle through the C ABI (identical LM traces, states within 1e-6 relative)."""
import numpy as np
import pytest

from oracle import oracle as orc
from sadvio_b200 import abi, api, synth

A = api.HUBER_A


def outlier_window(name="small", frac=0.04, seed=5):
    """A synth window with gross outliers in a few bearings (what the Huber loss is there for)."""
    win = synth.make_window(name)
    rng = np.random.default_rng(seed)
    bad = rng.choice(win.n_obs, max(3, int(frac * win.n_obs)), replace=False)
    b = win.obs_bearing[bad] + rng.normal(0, 0.02, (len(bad), 3))      # ~ 10 px at focal 458
    win.obs_bearing[bad] = b / np.linalg.norm(b, axis=1, keepdims=True)
    return win.normalise()


def rho(s, a):
    return np.where(s > a * a, 2 * a * np.sqrt(s) - a * a, s)


def robust_visual_cost(win, d, a):
    r, Jp, Jl, _ = orc.eval_visual(win, d)                              # raw functor output, no loss
    s = (r * r).sum(axis=1)
    return 0.5 * rho(s, a).sum() if a > 0 else 0.5 * s.sum(), r, Jp, Jl, s


def tight_cfg(iters=60):
    cfg = orc.default_config()
    cfg.function_tolerance = 1e-15
    cfg.parameter_tolerance = 1e-14
    cfg.gradient_tolerance = 1e-14
    cfg.max_num_iterations = iters
    return cfg


# ----------------------------------------------------------------------------------------------------------- oracle (CPU)
def test_landmark_window_layout():
    win = outlier_window()
    sub, lmks = api.landmark_window(win, 3)
    assert sub.n_fixed == sub.n_frames and not sub.vio and sub.max_num_iterations == 10 and sub.visual_loss_huber_a == A
    assert np.array_equal(lmks, np.unique(win.obs_lmk[win.obs_frame == 3]))
    # every observation of those landmarks, landmark-major, measurements carried over
    sel = np.isin(win.obs_lmk, lmks)
    assert sub.n_obs == int(sel.sum()) and np.all(np.diff(sub.obs_lmk) >= 0)
    assert np.array_equal(sub.obs_bearing, win.obs_bearing[sel]) and np.array_equal(lmks[sub.obs_lmk], win.obs_lmk[sel])


def test_landmark_optimization_oracle_cost_and_stationarity():
    win = outlier_window()
    sub, lmks = api.landmark_window(win, 3)
    rc, d, st = orc.solve_window(sub)
    assert rc == 0 and st["iterations"] <= 10 and st["final_cost"] < st["initial_cost"]
    assert np.all(d.dpose == 0)                                         # every pose block constant (…Analytic.cpp:141-145)
    c0, *_ = robust_visual_cost(sub, None, A)
    c1, *_ = robust_visual_cost(sub, d, A)
    assert abs(st["initial_cost"] - c0) <= 1e-12 * c0 and abs(st["final_cost"] - c1) <= 1e-12 * c1
    # the loss must be active on this data, and the cap of THIS window (10) must win over the configuration's 20
    _, r, _, _, s = robust_visual_cost(sub, None, A)
    assert (s > A * A).sum() > 10
    sub1 = api.landmark_window(win, 3)[0]
    sub1.max_num_iterations = 1
    assert orc.solve_window(sub1)[2]["iterations"] == 1
    # converged: stationary point of the ROBUST cost, gradient formed in numpy from the raw Jacobians (rho' = a / |r| beyond a)
    # (a few landmarks sit in the linear zone of the loss, where Gauss-Newton on the corrected residuals converges linearly)
    sub.max_num_iterations = 0
    rc, d, st = orc.solve_window(sub, tight_cfg(400))
    _, r, Jp, Jl, s = robust_visual_cost(sub, d, A)
    w = np.where(s > A * A, A / np.sqrt(np.maximum(s, 1e-300)), 1.0)
    g = np.zeros((sub.n_lmks, 3))
    np.add.at(g, sub.obs_lmk, w[:, None] * np.einsum("ork,or->ok", Jl.reshape(-1, 2, 3), r))
    scale = np.abs(w[:, None] * np.einsum("ork,or->ok", Jl.reshape(-1, 2, 3), r)).max()
    assert np.abs(g).max() < 1e-4 * scale


def test_huber_far_above_every_residual_is_no_loss():
    win = outlier_window()
    sub, _ = api.landmark_window(win, 2)
    sub.visual_loss_huber_a = 1e9
    rc, d, st = orc.solve_window(sub)
    sub.visual_loss_huber_a = 0.0
    rc0, d0, st0 = orc.solve_window(sub)
    assert st["iterations"] == st0["iterations"] and np.array_equal(d.dlmk, d0.dlmk)


def test_huber_resists_outliers():
    win = outlier_window(frac=0.08)
    sub, lmks = api.landmark_window(win, 3)
    gt = win.meta["lmk_gt"][lmks] if "lmk_gt" in win.meta else None
    if gt is None:
        pytest.skip("generator carries no landmark ground truth")
    sub.T_f_w = np.ascontiguousarray(win.meta["T_f_w_gt"][[int(np.flatnonzero((win.T_f_w == t).all(axis=1))[0]) for t in sub.T_f_w]])
    err = {}
    for a in (0.0, A):
        sub.visual_loss_huber_a = a
        sub.max_num_iterations = 0
        rc, d, st = orc.solve_window(sub, tight_cfg())
        err[a] = np.median(np.linalg.norm(sub.lmk_t + d.dlmk - gt, axis=1)), np.linalg.norm(sub.lmk_t + d.dlmk - gt, axis=1).max()
    assert err[A][1] < err[0.0][1]


@pytest.mark.parametrize("kind", [0, 1])
def test_single_frame_optimization_oracle(kind):
    win = synth.make_window("small", factor_kind=kind)
    sub, frames = api.single_frame_window(win, 0, vi=False)
    assert frames == [0] and sub.n_frames == 1 and sub.landmarks_constant and sub.max_num_iterations == 5 and not sub.vio
    assert (sub.obs_sigma is not None) == (kind == 0)
    rc, d, st = orc.solve_window(sub)
    assert rc == 0 and st["iterations"] <= 5 and st["final_cost"] < 0.5 * st["initial_cost"]
    assert np.all(d.dlmk == 0) and np.abs(d.dpose[0]).max() > 0        # landmarks constant (…Analytic.cpp:41-43)
    c1, *_ = robust_visual_cost(sub, d, 0.0)
    assert abs(st["final_cost"] - c1) <= 1e-10 * max(c1, 1e-30)
    # stationarity of the pose
    sub.max_num_iterations = 0
    rc, d, st = orc.solve_window(sub, tight_cfg())
    _, r, Jp, _, _ = robust_visual_cost(sub, d, 0.0)
    terms = np.einsum("ork,or->ok", Jp.reshape(-1, 2, 6), r)
    assert np.abs(terms.sum(axis=0)).max() < 1e-6 * np.abs(terms).max()
    if kind == 0:
        # the weight is 1 / sigma = focal, not the window solve's focal / 1.5 (…Analytic.cpp:48 vs :283)
        sub15 = api.single_frame_window(win, 0, vi=False)[0]
        sub15.obs_sigma = None
        c15 = orc.solve_window(sub15)[2]["initial_cost"]
        c10 = orc.solve_window(api.single_frame_window(win, 0, vi=False)[0])[2]["initial_cost"]
        assert abs(c10 / c15 - 2.25) < 1e-9


def test_single_frame_vi_optimization_oracle():
    win = synth.make_window("small")
    sub, frames = api.single_frame_window(win, 0, vi=True)
    assert frames == [0, int(win.imu_i[np.flatnonzero(win.imu_j == 0)[0]])] and sub.vio and sub.n_imu == 1 and sub.visual_loss_huber_a == A
    # both frames see the moving frame's landmarks; within a landmark the moving frame's features come first
    first = np.r_[True, sub.obs_lmk[1:] != sub.obs_lmk[:-1]]
    assert np.all(sub.obs_frame[first] == 0) and set(np.unique(sub.obs_frame)) == {0, 1}
    rc, d, st = orc.solve_window(sub)
    assert rc == 0 and st["iterations"] <= 5 and st["final_cost"] < st["initial_cost"]
    assert np.all(d.dlmk == 0) and np.abs(d.dpose[0]).max() > 0 and np.abs(d.dpose[1]).max() > 0 and np.abs(d.dv).max() > 0
    # the loss wraps the visual blocks only (the IMU factors are added with a null loss, AOptimizer.cpp:239)
    cv, *_ = robust_visual_cost(sub, None, A)
    r_imu, _, r_bias = orc.eval_imu(sub, None)
    assert abs(st["initial_cost"] - (cv + 0.5 * (r_imu ** 2).sum() + 0.5 * (r_bias ** 2).sum())) <= 1e-10 * st["initial_cost"]


def test_python_mirror_writes_back_like_the_reference():
    """landmarkOptimization updates only landmarks that pass sanityCheck (AOptimizer.cpp:132-140); the single-frame solves write
    the pose(s) back and nothing else (:199-203, :262-285).  The solve itself is replaced by the oracle here."""
    win = outlier_window()

    class OracleBacked(api.B200Optimizer):
        def __init__(self):
            self.last_stats = None

        def _solve(self, w):
            return orc.solve_window(w)

    opt = OracleBacked()
    before = win.lmk_t.copy()
    sub, lmks = api.landmark_window(win, 3)
    d = orc.solve_window(sub)[1]
    assert opt.landmarkOptimization(win, 3, sanity_check=lambda l: l % 2 == 0)
    moved = np.flatnonzero(np.abs(win.lmk_t - before).max(axis=1) > 0)
    assert set(moved) <= set(int(l) for l in lmks if l % 2 == 0) and len(moved) > 0
    k = int(np.flatnonzero(lmks == moved[0])[0])
    assert np.array_equal(win.lmk_t[moved[0]], before[moved[0]] + d.dlmk[k])
    T0, lm0, v0 = win.T_f_w.copy(), win.lmk_t.copy(), win.v.copy()
    assert opt.singleFrameOptimization(win, 1)
    assert np.array_equal(win.lmk_t, lm0) and np.array_equal(win.v, v0)
    assert np.abs(win.T_f_w[1] - T0[1]).max() > 0 and np.array_equal(np.delete(win.T_f_w, 1, 0), np.delete(T0, 1, 0))
    T0 = win.T_f_w.copy()
    dR0 = win.imu_dR.copy()
    assert opt.singleFrameVIOptimization(win, 0)
    i = int(win.imu_i[np.flatnonzero(win.imu_j == 0)[0]])
    changed = np.flatnonzero(np.abs(win.T_f_w - T0).max(axis=1) > 0)
    assert set(changed) == {0, i} and np.abs(win.v - v0).max() > 0
    assert np.array_equal(win.imu_dR, dR0)                              # no biasDeltaCorrection in the single-frame solves


# ----------------------------------------------------------------------------------------------------------- CUDA parity
def _parity(solver, sub, tol=1e-6):
    rc, d, st = solver.solve_window(sub)
    rc0, d0, st0 = orc.solve_window(sub)
    assert rc == rc0 == 0
    assert st["iterations"] == st0["iterations"] and st["termination"] == st0["termination"], (st["iterations"], st0["iterations"], st["termination"], st0["termination"])
    n = st["iterations"] + 1
    assert np.array_equal(np.asarray(st["trace_accepted"][:n]), np.asarray(st0["trace_accepted"][:n]))
    assert np.allclose(st["trace_cost"][:n], st0["trace_cost"][:n], rtol=1e-9, atol=0)
    assert abs(st["fixed_cost"] - st0["fixed_cost"]) <= 1e-9 * max(1.0, st0["fixed_cost"])
    for a, b in ((d.dpose, d0.dpose), (d.dv, d0.dv), (d.dba, d0.dba), (d.dbg, d0.dbg), (d.dlmk, d0.dlmk)):
        assert np.abs(a - b).max() <= tol * max(1e-12, np.abs(b).max()), (np.abs(a - b).max(), np.abs(b).max())
    return d, st


@pytest.mark.gpu
@pytest.mark.parametrize("name,frame", [("small", 3), ("C2", 5)])
def test_landmark_optimization_matches_oracle(solver, name, frame):
    win = outlier_window(name)
    sub, _ = api.landmark_window(win, frame)
    d, st = _parity(solver, sub)
    assert np.all(d.dpose == 0) and np.abs(d.dlmk).max() > 0 and st["iterations"] <= 10


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [0, 1])
def test_single_frame_optimization_matches_oracle(solver, kind):
    win = synth.make_window("C2", factor_kind=kind)
    for frame in (0, 7):
        sub, _ = api.single_frame_window(win, frame, vi=False)
        d, st = _parity(solver, sub)
        assert np.all(d.dlmk == 0) and np.abs(d.dpose).max() > 0 and st["iterations"] <= 5


@pytest.mark.gpu
def test_single_frame_vi_optimization_matches_oracle(solver):
    win = outlier_window("C2")
    sub, _ = api.single_frame_window(win, 0, vi=True)
    d, st = _parity(solver, sub)
    assert np.all(d.dlmk == 0) and np.abs(d.dv).max() > 0 and st["iterations"] <= 5


@pytest.mark.gpu
def test_huber_window_solve_matches_oracle(solver):
    """The loss is a field of the window, not of the three masks: a full VIO window with free poses AND landmarks under Huber."""
    win = outlier_window("small")
    win.visual_loss_huber_a = A
    _parity(solver, win)


@pytest.mark.gpu
def test_constant_landmarks_seen_from_a_fixed_keyframe_go_to_fixed_cost(solver):
    """landmarks_constant with a constant keyframe: its visual blocks have no free parameter block (Ceres fixed_cost)."""
    win = synth.make_window("small")
    sub, _ = api.single_frame_window(win, 0, vi=True)
    sub.n_fixed = 1                                                     # the previous keyframe (window index 1) is constant
    d, st = _parity(solver, sub)
    assert st["fixed_cost"] > 0 and np.all(d.dpose[1] == 0)


@pytest.mark.gpu
def test_f2_through_the_optimizer_interface(solver):
    win_a, win_b = outlier_window(), outlier_window()
    opt = api.B200Optimizer.__new__(api.B200Optimizer)
    opt.solver, opt.last_stats = solver, None

    class OracleBacked(api.B200Optimizer):
        def __init__(self):
            self.last_stats = None

        def _solve(self, w):
            return orc.solve_window(w)

    ref = OracleBacked()
    for o, w in ((opt, win_a), (ref, win_b)):
        assert o.landmarkOptimization(w, 3) and o.singleFrameOptimization(w, 1) and o.singleFrameVIOptimization(w, 0)
    for a, b in ((win_a.T_f_w, win_b.T_f_w), (win_a.lmk_t, win_b.lmk_t), (win_a.v, win_b.v), (win_a.ba, win_b.ba), (win_a.bg, win_b.bg)):
        assert np.abs(a - b).max() <= 1e-6 * np.abs(b).max()

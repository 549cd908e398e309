"""AOptimizer::VIInit (reference cpp/src/optimizers/AOptimizer.cpp:448-581, functor IMUFactorInit residuals.hpp:302-410): the
fourth "other" solve of SURVEY.md section 8 f2 — gravity direction, keyframe velocities and (optionally) the metric scale of an
up-to-scale trajectory from the IMU pre-integration alone.

CPU tests pin the oracle's restatement: the functor's Jacobians against central differences (they are exact derivatives except
the one with respect to the log-scale, which omits exp(lambda) in the reference — reproduced and tested as such), and the
reference's own end-to-end test (imu_test.cpp:813-880: 10 EuRoC keyframes scaled by 0.5 must come back onto the ground truth
within 0.02).  GPU tests compare sdv_viinit (one kernel, sdv_viinit.cuh) with the oracle through the C ABI: same LM trace,
parameter blocks within 1e-6 relative."""
import copy

import numpy as np
import pytest

from oracle import oracle as orc
from sadvio_b200 import api, synth
from tests import ref_fixtures as rf


def _pre_of(win, p):
    return orc.pack_preint(win.imu_dR[p], win.imu_dv[p], win.imu_dp[p], win.imu_cov[p], win.imu_J_dR_bg[p], win.imu_J_dv_ba[p],
                           win.imu_J_dv_bg[p], win.imu_J_dp_ba[p], win.imu_J_dp_bg[p])


def _eval(win, p, par, jac=True):
    i, j = int(win.imu_i[p]), int(win.imu_j[p])
    return orc.imu_init_eval(win.T_f_w[i], win.T_f_w[j], win.v[i], win.v[j], win.imu_dt[p], _pre_of(win, p), par, jac)


def rotated_scaled(win, rot, scale):
    """The same trajectory expressed in a world frame rotated by Exp(rot) and shrunk by `scale` — what a monocular front end
    hands to VIInit.  T_f_w <- [R_f_w R_g^T | scale t_f_w], v <- scale R_g v."""
    Rg = orc.exp_so3(rot)
    w = copy.deepcopy(win)
    for f in range(w.n_frames):
        T = w.T_f_w[f].reshape(3, 4).copy()
        T[:, :3] = T[:, :3] @ Rg.T
        T[:, 3] *= scale
        w.T_f_w[f] = T.reshape(12)
        w.v[f] = scale * (Rg @ w.v[f])
    w.lmk_t = scale * (w.lmk_t @ Rg.T)
    return w.normalise()


# ----------------------------------------------------------------------------------------------------------- oracle (CPU)
def test_imu_init_factor_jacobians_match_central_differences():
    win, _ = rf.euroc_viinit_window()
    rng = np.random.default_rng(3)
    for p in (0, 4, 8):
        par = np.concatenate([rng.normal(0, 0.05, 2), rng.normal(0, 0.1, 6), rng.normal(0, 0.01, 3), rng.normal(0, 0.005, 3), [0.0]])
        r, J = _eval(win, p, par)
        num = np.zeros((9, 15))
        for k in range(15):
            h = 1e-6
            a, b = par.copy(), par.copy()
            a[k] += h
            b[k] -= h
            num[:, k] = (_eval(win, p, a, False)[0] - _eval(win, p, b, False)[0]) / (2 * h)
        scale = np.abs(num).max()
        assert np.abs(J - num).max() <= 2e-7 * scale, (p, np.abs(J - num).max(), scale)


def test_imu_init_factor_scale_jacobian_omits_exp_lambda():
    # residuals.hpp:396-402: d r_dp / d lambda is written without the exp(lambda) the residual carries (:333); Ceres uses it as is
    win, _ = rf.euroc_viinit_window()
    par = np.zeros(15)
    par[14] = 0.4
    r, J = _eval(win, 2, par)
    h = 1e-6
    a, b = par.copy(), par.copy()
    a[14] += h
    b[14] -= h
    num = (_eval(win, 2, a, False)[0] - _eval(win, 2, b, False)[0]) / (2 * h)
    assert np.abs(J[:, 14] * np.exp(0.4) - num).max() <= 1e-7 * np.abs(num).max()
    assert np.abs(J[:, 14] - num).max() > 0.1 * np.abs(num).max()


def test_reference_imu_factor_init_solve():
    """imu_test.cpp:489-541: after the free-fall inertial optimisation both poses are shrunk by 0.5 and ONE IMUFactorInit with all six
    parameter blocks free (gravity direction, both velocities, dba, dbg, log-scale) is solved with Ceres' defaults (50 iterations):
    the recovered scale must undo the shrinking, |0.5 - 1 / exp(lambda)| < 1e-2."""
    win = rf.free_fall_window()
    rc, d, st = orc.solve_window(win)                            # :473-483 (checked by tests/test_oracle_solver.py)
    assert rc == 0
    api.write_back(win, d, True)
    for f in range(2):                                           # :490-496
        T = win.T_f_w[f].reshape(3, 4).copy()
        T[:, 3] *= 0.5
        win.T_f_w[f] = T.reshape(12)
    win.max_num_iterations = 50                                  # ceres::Solver::Options default
    rc, res, st = orc.viinit(win, True, all_blocks_free=True)
    assert rc == 0 and st["final_cost"] < 1e-3 * st["initial_cost"]
    assert abs(0.5 - 1.0 / np.exp(res["lam"])) < 1e-2            # :541


def test_oracle_viinit_reference_euroc_run():
    """imu_test.cpp:813-880 through the oracle + the host write-back (AOptimizer.cpp:531-567)."""
    win, gt = rf.euroc_viinit_window()
    before = max(np.linalg.norm(gt[f] @ np.vstack([win.T_f_w[f].reshape(3, 4), [0, 0, 0, 1]]) - np.eye(4)) for f in range(win.n_frames))
    assert before > 0.5                                          # the scaled trajectory is far from the ground truth
    rc, res, st = orc.viinit(win, True)
    assert rc == 0 and st["iterations"] <= 50
    assert abs(res["scale"] - 2.0) < 0.02                        # scale_factor = 0.5 recovered (:817)
    assert np.abs(res["R_w_i"] - np.eye(3)).max() < 1e-3         # the EuRoC world frame is already gravity-aligned
    api.viinit_write_back(win, res)
    rf.check_euroc_viinit(win, gt)


def _golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "viinit_euroc.npz"))


def test_oracle_viinit_matches_the_committed_golden():
    """tests/golden/viinit_euroc.npz (tests/golden/make_golden.py) pins the oracle's VIInit against drift."""
    g = _golden()
    win, _ = rf.euroc_viinit_window()
    rc, res, st = orc.viinit(win, True)
    assert rc == 0 and st["iterations"] == int(g["iterations"]) and st["termination"] == str(g["termination"])
    n = st["iterations"] + 1
    assert np.allclose(st["trace_cost"][:n], g["trace_cost"][:n], rtol=1e-10) and np.array_equal(np.asarray(st["trace_accepted"][:n]), g["trace_accepted"][:n])
    assert np.abs(res["dv"] - g["dv"]).max() <= 1e-9 * np.abs(g["dv"]).max() and abs(res["lam"] - float(g["lam"])) <= 1e-9
    assert np.abs(res["r_wi"] - g["r_wi"]).max() <= 1e-9 * max(np.abs(g["r_wi"]).max(), 1e-6)


def test_oracle_viinit_recovers_gravity_direction_with_fixed_scale():
    win, _ = rf.euroc_viinit_window(scale_factor=1.0)
    rot = np.array([0.12, -0.2, 0.0])
    w = rotated_scaled(win, rot, 1.0)
    rc, res, st = orc.viinit(w, False)
    assert rc == 0 and res["lam"] == 0.0 and res["scale"] == 1.0
    # the world frame of `w` is the inertial one rotated by Exp(rot): R_w_i maps inertial -> w, up to the yaw VIInit cannot see
    g_w = res["R_w_i"] @ np.array([0, 0, -1.0])
    assert np.linalg.norm(g_w - orc.exp_so3(rot) @ np.array([0, 0, -1.0])) < 2e-3
    assert st["final_cost"] < 1e-3 * st["initial_cost"]


def test_viinit_write_back_rules():
    win, _ = rf.euroc_viinit_window()
    win.has_prior = np.array([0] * (win.n_frames - 1) + [1], np.uint8)
    win.lmk_t = np.array([[1.0, 2.0, 3.0], [-1.0, 0.5, 2.0]])
    T0, v0, l0 = win.T_f_w.copy(), win.v.copy(), win.lmk_t.copy()
    Rwi = orc.exp_so3([0.1, -0.05, 0.0])
    res = dict(dv=np.full((win.n_frames, 3), 0.25), R_w_i=Rwi, lam=np.log(2.0))
    api.viinit_write_back(win, res)
    assert np.allclose(win.v, v0 + 0.25)
    for f in range(win.n_frames):
        T, Tn = T0[f].reshape(3, 4), win.T_f_w[f].reshape(3, 4)
        assert np.allclose(Tn[:, :3], T[:, :3] @ Rwi) and np.allclose(Tn[:, 3], 2.0 * T[:, 3])
    assert np.array_equal(win.T_prior[-1], win.T_f_w[-1]) and np.all(win.inf_prior[-1] == 100.0)
    assert np.allclose(win.lmk_t, 2.0 * (l0 @ Rwi))           # exp(lambda) R_w_i^T t  (AOptimizer.cpp:562-563)
    # a landmark seen from a frame keeps its frame coordinates up to the scale: R_f_w t_lmk + t_f_w
    for f in (0, 5):
        a = T0[f].reshape(3, 4)[:, :3] @ l0[0] + T0[f].reshape(3, 4)[:, 3]
        b = win.T_f_w[f].reshape(3, 4)[:, :3] @ win.lmk_t[0] + win.T_f_w[f].reshape(3, 4)[:, 3]
        assert np.allclose(b, 2.0 * a)


# ------------------------------------------------------------------------------------------------------------- CUDA (GPU)
def _assert_same(res, st, res0, st0, tol=1e-6):
    assert st["iterations"] == st0["iterations"] and st["termination"] == st0["termination"], (st["iterations"], st0["iterations"], st["termination"], st0["termination"])
    n = st0["iterations"] + 1
    assert np.array_equal(np.asarray(st["trace_accepted"][:n]), np.asarray(st0["trace_accepted"][:n]))
    assert np.allclose(st["trace_cost"][:n], st0["trace_cost"][:n], rtol=1e-7, atol=1e-12)
    for k in ("dv", "r_wi"):
        assert np.abs(res[k] - res0[k]).max() <= tol * max(np.abs(res0[k]).max(), 1e-9), (k, np.abs(res[k] - res0[k]).max())
    assert abs(res["lam"] - res0["lam"]) <= tol * max(abs(res0["lam"]), 1e-9)
    assert np.abs(res["R_w_i"] - res0["R_w_i"]).max() <= tol and abs(res["scale"] - res0["scale"]) <= tol * res0["scale"]


@pytest.mark.gpu
def test_viinit_reference_euroc_run_on_the_gpu(solver):
    win, gt = rf.euroc_viinit_window()
    rc, res, st = solver.viinit(win, True)
    rc0, res0, st0 = orc.viinit(win, True)
    assert rc == rc0 == 0
    _assert_same(res, st, res0, st0)
    assert st["kernel_launches"] == 2 and st["n_reduced"] == 3 * win.n_frames + 3
    api.viinit_write_back(win, res)
    rf.check_euroc_viinit(win, gt)                                # imu_test.cpp:873-878


@pytest.mark.gpu
def test_viinit_matches_the_committed_golden(solver):
    """sdv_viinit against tests/golden/viinit_euroc.npz: same 47-step trace, parameter blocks within 1e-6 relative."""
    g = _golden()
    win, _ = rf.euroc_viinit_window()
    rc, res, st = solver.viinit(win, True)
    assert rc == 0 and st["iterations"] == int(g["iterations"]) and st["termination"] == str(g["termination"])
    n = st["iterations"] + 1
    assert np.array_equal(np.asarray(st["trace_accepted"][:n]), g["trace_accepted"][:n])
    assert np.allclose(st["trace_cost"][:n], g["trace_cost"][:n], rtol=1e-7, atol=1e-12)
    assert np.abs(res["dv"] - g["dv"]).max() <= 1e-6 * np.abs(g["dv"]).max()
    assert abs(res["lam"] - float(g["lam"])) <= 1e-6 * abs(float(g["lam"])) and abs(res["scale"] - float(g["scale"])) <= 1e-6 * float(g["scale"])
    assert np.abs(res["r_wi"] - g["r_wi"]).max() <= 1e-6 * max(np.abs(g["r_wi"]).max(), 1e-9) + 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("optim_scale", [False, True])
def test_viinit_rotated_world_matches_oracle(solver, optim_scale):
    win, _ = rf.euroc_viinit_window(scale_factor=1.0)
    w = rotated_scaled(win, np.array([0.12, -0.2, 0.0]), 0.7 if optim_scale else 1.0)
    rng = np.random.default_rng(11)
    w.v += rng.normal(0, 0.05, w.v.shape)
    rc, res, st = solver.viinit(w, optim_scale)
    rc0, res0, st0 = orc.viinit(w, optim_scale)
    assert rc == rc0 == 0
    _assert_same(res, st, res0, st0)


@pytest.mark.gpu
def test_viinit_fifty_keyframes_matches_oracle(solver):
    # the C3 window's 50 keyframes / 49 IMU pairs (n = 153: the largest system that stays in shared memory), world rotated and shrunk
    win = synth.make_window("C3")
    w = rotated_scaled(win, np.array([-0.08, 0.15, 0.0]), 0.8)
    rc, res, st = solver.viinit(w, True)
    rc0, res0, st0 = orc.viinit(w, True)
    assert rc == rc0 == 0 and st["n_reduced"] == 153
    _assert_same(res, st, res0, st0)


@pytest.mark.gpu
def test_viinit_sixty_keyframes_normal_equations_in_global_memory(solver):
    # n = 3 * 60 + 3 = 183 > 158: the normal equations no longer fit the 200 KB of shared memory, the kernel works in its global scratch
    win = synth.make_window(synth.SynthConfig("vi60", 60, 600, span=4, seed=20260925 + 160))
    w = rotated_scaled(win, np.array([0.05, -0.1, 0.0]), 0.9)
    rc, res, st = solver.viinit(w, True)
    rc0, res0, st0 = orc.viinit(w, True)
    assert rc == rc0 == 0 and st["n_reduced"] == 183
    _assert_same(res, st, res0, st0)


@pytest.mark.gpu
def test_viinit_through_the_optimizer_mirror(solver):
    win, gt = rf.euroc_viinit_window()
    opt = api.B200Optimizer.__new__(api.B200Optimizer)
    opt.solver, opt.last_stats = solver, None
    scale, R_w_i = opt.VIInit(win, True)
    assert abs(scale - 2.0) < 0.02 and R_w_i.shape == (3, 3)
    rf.check_euroc_viinit(win, gt)


@pytest.mark.gpu
def test_viinit_iteration_cap_and_frame_without_pair(solver):
    win, _ = rf.euroc_viinit_window()
    win.max_num_iterations = 7
    # drop the newest pair: frame 0 keeps a velocity block no factor touches (Ceres never sees it; its update must stay zero)
    for name in ("imu_i", "imu_j", "imu_dt", "imu_dR", "imu_dv", "imu_dp", "imu_cov", "imu_J_dR_bg", "imu_J_dv_ba", "imu_J_dv_bg",
                 "imu_J_dp_ba", "imu_J_dp_bg", "imu_sigma_ba", "imu_sigma_bg"):
        setattr(win, name, getattr(win, name)[1:])
    rc, res, st = solver.viinit(win, True)
    rc0, res0, st0 = orc.viinit(win, True)
    assert st["iterations"] == st0["iterations"] == 7 and st["termination"] == st0["termination"] == "NO_CONVERGENCE"
    _assert_same(res, st, res0, st0)
    assert np.all(res["dv"][0] == 0.0)

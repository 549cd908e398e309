"""The oracle's restated Ceres LM loop: pinned through the reference's own end-to-end tests and self-consistency."""
import numpy as np
import pytest

from oracle import oracle as orc
from sadvio_b200 import abi, synth
from tests import ref_fixtures as rf


def test_reference_inertial_optimisation():  # imu_test.cpp:464-487 (tolerances 1e-2 / 1e-5)
    win = rf.free_fall_window()
    for mode in (0, 1):
        rc, d, st = orc.solve_window(win, mode=mode)
        assert rc == 0
        rf.check_free_fall(win, d)
        assert st["final_cost"] < st["initial_cost"]


def test_reference_bias_estimation():  # imu_test.cpp:545-568 (tolerance 1e-5)
    win = rf.bias_window()
    rc, d, st = orc.solve_window(win, mode=0)
    assert rc == 0
    rf.check_bias(win, d)


def test_c1_euroc_plumbing_window():
    """BASELINE config 1 (CPU only): 10 keyframes 0.5 s apart on the EuRoC ground truth from sample 2000 on, IMU
    by differentiating it as imu_test.cpp:741-757 does, the stereo rig of eth.yaml, ~10 k observations."""
    win = synth.make_window("C1")
    assert (win.n_frames, win.n_lmks, win.n_obs, win.n_imu) == (10, 1250, 10000, 9) and win.vio and win.n_fixed == 1
    assert np.allclose(win.imu_dt, 0.5)                              # < 1 s: no IMU factor is skipped (AOptimizer.cpp:69)
    # the pre-integrated deltas of the generator equal the oracle's processIMU chain on this trajectory too
    ts, R, p_gt, v_gt = synth.load_euroc_slice()
    gt = win.meta["T_f_w_gt"]
    R0 = gt[-1].reshape(3, 4)[:, :3].T                               # oldest keyframe = EuRoC sample 2000
    assert np.abs(R0 - R[0]).max() < 1e-12
    rc, d, st = orc.solve_window(win, nthreads=4)
    assert rc == 0 and st["termination"] in ("FUNCTION_TOLERANCE", "PARAMETER_TOLERANCE") and st["final_cost"] < 0.01 * st["initial_cost"]
    new = synth.apply_delta(win, d)
    before, after = np.abs(win.T_f_w - gt).max(), np.abs(new["T_f_w"] - gt).max()
    assert after < 0.1 * before
    assert np.abs(new["v"] - win.meta["v_gt"]).max() < 0.05
    # Schur elimination == the full normal equations SPARSE_NORMAL_CHOLESKY factors (the reference's solver for this config)
    rc1, d1, st1 = orc.solve_window(win, mode=1)
    assert st1["iterations"] == st["iterations"] and np.abs(d1.dpose - d.dpose).max() < 1e-9 * max(1.0, np.abs(d.dpose).max())


def test_keyframe_without_imu_in_a_vio_window():
    """A keyframe whose getIMU() is null gets no velocity / bias blocks and no IMU factor (AOptimizer.cpp:30-52, 60-73).  On the
    flattened window its v / ba / bg columns simply carry no factor: their gradient and Hessian are zero, the LM damping keeps
    them at zero, and norms / costs are what they are without the blocks — so `has_imu` needs no special path."""
    win = synth.make_window("small")
    f = 4
    keep = (win.imu_i != f) & (win.imu_j != f)
    for k in ("imu_i", "imu_j", "imu_dt", "imu_dR", "imu_dv", "imu_dp", "imu_cov", "imu_J_dR_bg", "imu_J_dv_ba", "imu_J_dv_bg", "imu_J_dp_ba",
              "imu_J_dp_bg", "imu_sigma_ba", "imu_sigma_bg"):
        setattr(win, k, getattr(win, k)[keep].copy())
    win.has_imu[f] = 0
    rc, d, st = orc.solve_window(win.normalise())
    assert rc == 0 and st["final_cost"] < 0.01 * st["initial_cost"]
    assert np.all(d.dv[f] == 0) and np.all(d.dba[f] == 0) and np.all(d.dbg[f] == 0) and np.abs(d.dpose[f]).max() > 0
    rc1, d1, st1 = orc.solve_window(win, mode=1)                        # full normal equations: same step
    assert st1["iterations"] == st["iterations"] and np.abs(d1.dpose - d.dpose).max() < 1e-9


def test_reference_euroc_bias_run():  # imu_test.cpp:885-945 (tolerance 0.02 on both biases of the 30th keyframe)
    """The reference's longest end-to-end test of the window solve: 29 consecutive localMapVIOptimization calls on a growing
    window, each followed by the state write-back and biasDeltaCorrection, starting from zero biases."""
    kfs, stats = rf.euroc_bias_run(rf.OracleOptimizer())
    rf.check_euroc_bias(kfs)
    assert len(stats) == 29 and all(s["termination"] in ("FUNCTION_TOLERANCE", "PARAMETER_TOLERANCE", "GRADIENT_TOLERANCE") for s in stats)
    # the fixed oldest keyframe never moves (AOptimizer.cpp:46-51)
    assert np.all(orc.imu_get(kfs[0]["imu"], "ba") == 0) and np.all(orc.imu_get(kfs[0]["imu"], "bg") == 0)


@pytest.mark.parametrize("name,kind", [("tiny", 0), ("tiny", 1), ("small", 0), ("small", 1)])
def test_schur_equals_full_normal_equations(name, kind):
    """Landmark elimination must give the step SPARSE_NORMAL_CHOLESKY computes on the full system."""
    win = synth.make_window(name, factor_kind=kind)
    rc0, d0, st0 = orc.solve_window(win, mode=0)
    rc1, d1, st1 = orc.solve_window(win, mode=1)
    assert rc0 == rc1 == 0
    assert st0["iterations"] == st1["iterations"] and st0["termination"] == st1["termination"]
    assert st0["trace_accepted"] == st1["trace_accepted"]
    for a, b in ((d0.dpose, d1.dpose), (d0.dv, d1.dv), (d0.dba, d1.dba), (d0.dbg, d1.dbg), (d0.dlmk, d1.dlmk)):
        assert np.abs(a - b).max() <= 1e-9 * max(1.0, np.abs(b).max())


def test_converges_to_ground_truth():
    win = synth.make_window("small")
    cfg = orc.default_config()
    rc, d, st = orc.solve_window(win, cfg)
    new = synth.apply_delta(win, d)
    gt = win.meta
    assert st["termination"] == "FUNCTION_TOLERANCE" and st["iterations"] <= 20
    assert np.abs(new["T_f_w"] - gt["T_f_w_gt"]).max() < 0.02 < np.abs(win.T_f_w - gt["T_f_w_gt"]).max()
    assert np.abs(new["bg"] - gt["bg_true"]).max() < 1e-4
    assert np.abs(new["ba"] - gt["ba_true"]).max() < 5e-3


def test_threads_do_not_change_the_answer():
    win = synth.make_window("small")
    _, d1, s1 = orc.solve_window(win, nthreads=1)
    _, d4, s4 = orc.solve_window(win, nthreads=4)
    assert s1["iterations"] == s4["iterations"]
    assert np.abs(d1.dpose - d4.dpose).max() < 1e-10


def test_fixed_cost_and_constant_blocks():
    """A pose prior on a fixed keyframe has only constant blocks: Ceres moves it to fixed_cost."""
    win = synth.make_window("tiny")
    # move the fixed frame's prior away from its pose so that the prior has a non-zero residual
    win.T_prior[-1] = synth.T34(np.vstack([win.T_f_w[-1].reshape(3, 4), [0, 0, 0, 1]]) @ np.diag([1, 1, 1, 1.0]) +
                                np.array([[0, 0, 0, 0.01]] * 3 + [[0, 0, 0, 0]]))
    rc, d, st = orc.solve_window(win)
    assert st["fixed_cost"] > 0
    assert np.all(d.dpose[-1] == 0) and np.all(d.dv[-1] == 0) and np.all(d.dba[-1] == 0)
    # with no fixed frame the same prior is part of the cost
    win.n_fixed = 0
    rc, d, st2 = orc.solve_window(win)
    assert st2["fixed_cost"] == 0


def test_pre_integration_generator_matches_oracle():
    """synth.preintegrate (numpy, product side) == oracle processIMU chain, including the stale delta_R quirk."""
    rng = np.random.default_rng(7)
    n, dt = 20, 0.005
    acc = rng.normal(0, 1, (2 * n, 3)) + [0, 0, 9.81]
    gyr = rng.normal(0, 0.3, (2 * n, 3))
    ba, bg = np.array([0.02, -0.01, 0.03]), np.array([0.001, 0.002, -0.001])
    e = rf.eta(200.0)
    p1 = synth.preintegrate(acc[:n], gyr[:n], dt, ba, bg, e, None)
    p2 = synth.preintegrate(acc[n:], gyr[n:], dt, ba, bg, e, p1)
    last = orc.imu_state(acc[0], gyr[0], ba=ba, bg=bg, is_kf=True)
    for k in range(1, 2 * n + 1):
        a, g = (acc[k], gyr[k]) if k < 2 * n else (acc[-1], gyr[-1])
        last = orc.process_imu(last, ba, bg, dt, e, 200.0, a, g, is_kf=(k == n))
        if k == n:
            kf1 = last
    for pre, st in ((p1, kf1), (p2, last)):
        assert np.allclose(pre.dR.reshape(9), orc.imu_get(st, "dR"), atol=1e-14)
        assert np.allclose(pre.dv, orc.imu_get(st, "dv"), atol=1e-14)
        assert np.allclose(pre.dp, orc.imu_get(st, "dp"), atol=1e-14)
        assert np.allclose(pre.cov.reshape(81), orc.imu_get(st, "Sigma"), rtol=1e-12, atol=1e-22)
        for nm in ("J_dR_bg", "J_dv_ba", "J_dv_bg", "J_dp_ba", "J_dp_bg"):
            assert np.allclose(getattr(pre, nm).reshape(9), orc.imu_get(st, nm), atol=1e-14)


@pytest.mark.parametrize("vio", [True, False])
def test_dense_marginal_prior_schur_equals_full(vio):
    """MarginalizationFactor couples kept landmarks: they stay in the reduced system (elimination group 2)."""
    win = synth.add_dense_prior(synth.make_window("small", vio=vio), n_keep=15, with_frame=vio)
    rc0, d0, st0 = orc.solve_window(win, mode=0)
    rc1, d1, st1 = orc.solve_window(win, mode=1)
    assert rc0 == rc1 == 0 and st0["iterations"] == st1["iterations"]
    n_cols = 15 * 9 if vio else 6 * 9
    n_kept = int(np.sum(win.dense_prior.keep_col >= 0))
    assert st0["n_reduced"] == n_cols + 3 * n_kept
    assert np.abs(d0.dpose - d1.dpose).max() < 1e-9 and np.abs(d0.dlmk - d1.dlmk).max() < 1e-8
    # the prior changes the solution
    _, d2, _ = orc.solve_window(synth.make_window("small", vio=vio))
    assert np.abs(d0.dpose - d2.dpose).max() > 1e-6


def test_sparsified_priors_schur_equals_full():
    """a11: IMUPriordx + PoseToLandmark (VIO) and Landmark3DPrior + LandmarkToLandmark chain (VO)."""
    for win in (synth.add_sparse_prior_vio(synth.make_window("small"), 15), synth.add_sparse_prior_vo(synth.make_window("small", vio=False), 12)):
        rc0, d0, st0 = orc.solve_window(win, mode=0)
        rc1, d1, st1 = orc.solve_window(win, mode=1)
        assert rc0 == rc1 == 0 and st0["iterations"] == st1["iterations"]
        assert np.abs(d0.dpose - d1.dpose).max() < 1e-9 and np.abs(d0.dlmk - d1.dlmk).max() < 1e-8
    # chain landmarks live in the reduced system
    assert st0["n_reduced"] == 6 * 9 + 3 * 12


# ---------------------------------------------------------------------------------------------------------------------
# committed golden solutions (tests/golden/make_golden.py): the oracle must keep reproducing them
# ---------------------------------------------------------------------------------------------------------------------
import os  # noqa: E402

import pytest  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = {"tiny_vio_angular": ("tiny", dict(factor_kind=0)), "tiny_vio_pixel": ("tiny", dict(factor_kind=1)),
                "small_vio_angular": ("small", dict(factor_kind=0))}


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_oracle_reproduces_golden_solution(name):
    import numpy as np

    from oracle import oracle
    from sadvio_b200 import synth

    cfg, kw = GOLDEN_CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    win = synth.make_window(cfg, **kw)
    # landmark-keyframe visibility indices: bit-exact
    assert np.array_equal(win.obs_lmk, g["obs_lmk"]) and np.array_equal(win.obs_frame, g["obs_frame"]) and np.array_equal(win.obs_cam, g["obs_cam"])
    rc, d, st = oracle.solve_window(win, mode=0, nthreads=4)
    assert rc == 0 and st["iterations"] == int(g["iterations"]) and st["termination"] == str(g["termination"])
    assert list(st["trace_accepted"]) == list(g["trace_accepted"])
    for k in ("dpose", "dv", "dba", "dbg", "dlmk"):
        a, b = getattr(d, k), g[k]
        assert np.abs(a - b).max() <= 1e-9 * max(1e-12, np.abs(b).max())

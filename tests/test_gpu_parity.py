"""GPU parity tests (run with -m gpu on a B200). Everything goes through the C ABI (ctypes -> libsadvio_b200.so).

Bars: landmark-keyframe visibility indices bit-exact; residuals/Jacobians <= 1e-12 relative per element;
pose / velocity / bias states <= 1e-6 relative after the full LM solve (BASELINE.json north_star)."""
import numpy as np
import pytest

from oracle import oracle as orc
from sadvio_b200 import abi, api, synth
from tests import ref_fixtures as rf

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(1e-300, np.abs(b).max()))


def random_point(win, seed):
    rng = np.random.default_rng(seed)
    F, L = win.n_frames, win.n_lmks
    x = abi.Delta(rng.normal(0, 0.02, (F, 6)), rng.normal(0, 0.02, (F, 3)), rng.normal(0, 0.01, (F, 3)),
                  rng.normal(0, 0.001, (F, 3)), rng.normal(0, 0.05, (L, 3)))
    nf = win.n_fixed
    if nf:
        x.dpose[-nf:] = 0
        x.dv[-nf:] = 0
        x.dba[-nf:] = 0
        x.dbg[-nf:] = 0
    if not win.vio:
        x.dv[:] = 0
        x.dba[:] = 0
        x.dbg[:] = 0
    return x


@pytest.mark.parametrize("name,kind", [("tiny", 0), ("tiny", 1), ("small", 0), ("small", 1), ("C2", 0), ("C2", 1)])
def test_visual_residuals_and_jacobians(solver, name, kind):
    win = synth.make_window(name, factor_kind=kind)
    solver.upload(win)
    for x in (None, random_point(win, 1)):
        r, Jp, Jl, _ = solver.eval_visual(x)
        r0, Jp0, Jl0, _ = orc.eval_visual(win, x)
        assert rel(r, r0) < 1e-12 and rel(Jp, Jp0) < 1e-12 and rel(Jl, Jl0) < 1e-12


@pytest.mark.parametrize("name", ["tiny", "small", "C2"])
def test_imu_factors(solver, name):
    win = synth.make_window(name)
    solver.upload(win)
    for x in (None, random_point(win, 2)):
        r, J, rb = solver.eval_imu(x)
        r0, J0, rb0 = orc.eval_imu(win, x)
        assert rel(r, r0) < 1e-10 and rel(J, J0) < 1e-10 and rel(rb, rb0) < 1e-12


def test_pixel_failed_projection_branch(solver):
    """Residual zeroed, Jacobian kept when the projection fails (BundleAdjustmentCERESAnalytic.h:63-68)."""
    win = synth.make_window("tiny", factor_kind=1)
    win.lmk_t[0] += 50.0       # throws landmark 0 out of every image
    win.lmk_t[1] = -win.lmk_t[1]  # behind the cameras
    solver.upload(win)
    r, Jp, Jl, _ = solver.eval_visual(None)
    r0, Jp0, Jl0, _ = orc.eval_visual(win, None)
    m = win.obs_lmk <= 1
    assert np.all(r0[m] == 0.0) and np.all(r[m] == 0.0)
    assert np.abs(Jl[m]).max() > 0
    assert rel(Jp, Jp0) < 1e-12 and rel(Jl, Jl0) < 1e-12 and rel(r, r0) < 1e-12


def assert_same_states(win, d, d0, tol=1e-6):
    """The bar as BASELINE.json words it: updated pose / velocity / bias STATES agree to 1e-6 relative."""
    a, b = synth.apply_delta(win, d), synth.apply_delta(win, d0)
    keys = ("T_f_w", "v", "ba", "bg") if win.vio else ("T_f_w",)
    for k in keys:
        assert np.abs(a[k] - b[k]).max() <= tol * np.abs(b[k]).max()


def solve_both(solver, win, cfg=None):
    if cfg is not None:
        s = api.Solver(cfg)
    else:
        s = solver
    rc, d, st = s.solve_window(win)
    rc0, d0, st0 = orc.solve_window(win, cfg, mode=0, nthreads=8)
    if cfg is not None:
        s.close()
    return (rc, d, st), (rc0, d0, st0)


def assert_same_solution(g, o, tol=1e-6):
    (rc, d, st), (rc0, d0, st0) = g, o
    assert rc == rc0
    assert st["iterations"] == st0["iterations"], (st["trace_cost"], st0["trace_cost"])
    assert st["termination"] == st0["termination"]
    assert st["trace_accepted"] == st0["trace_accepted"]
    assert st["num_successful_steps"] == st0["num_successful_steps"]
    assert rel(st["trace_cost"], st0["trace_cost"]) < 1e-9
    assert rel(st["trace_radius"], st0["trace_radius"]) < 1e-6
    assert abs(st["fixed_cost"] - st0["fixed_cost"]) <= 1e-9 * max(1.0, st0["fixed_cost"])
    # north_star bar: 1e-6 relative on pose / velocity / bias states (blocks that did not move are compared absolutely)
    for a, b in ((d.dpose, d0.dpose), (d.dv, d0.dv), (d.dba, d0.dba), (d.dbg, d0.dbg)):
        assert np.abs(a - b).max() <= tol * np.abs(b).max() + 1e-13
    if d0.dlmk.size:
        assert np.abs(d.dlmk - d0.dlmk).max() <= 1e-5 * np.abs(d0.dlmk).max() + 1e-13


@pytest.mark.parametrize("name,kind", [("tiny", 0), ("tiny", 1), ("small", 0), ("small", 1), ("C2", 0), ("C2", 1)])
def test_full_solve_matches_oracle(solver, name, kind):
    win = synth.make_window(name, factor_kind=kind)
    g, o = solve_both(solver, win)
    assert_same_solution(g, o)
    assert_same_states(win, g[1], o[1])


def test_full_solve_c3_headline(solver):
    """BASELINE.json config 3: 50 KF x 10k landmarks x 80k observations, full VIO factor set."""
    win = synth.make_window("C3")
    assert (win.n_frames, win.n_lmks, win.n_obs, win.n_imu) == (50, 10000, 80000, 49)
    g, o = solve_both(solver, win)
    assert_same_solution(g, o)
    new = synth.apply_delta(win, g[1])
    gt = win.meta
    assert np.abs(new["T_f_w"] - gt["T_f_w_gt"]).max() < 0.02


def test_tight_convergence_matches_oracle(solver):
    """Trajectory effects vanish when both sides converge tightly (SURVEY §7 'hard parts')."""
    win = synth.make_window("small")
    cfg = api.default_config()
    cfg.function_tolerance = 1e-12
    cfg.max_num_iterations = 50
    g, o = solve_both(solver, win, cfg)
    assert_same_solution(g, o, tol=1e-7)


def test_ba_mode_no_imu(solver):
    """localMapBA: 6 dof per frame, no IMU blocks (AOptimizer.cpp:299-350)."""
    win = synth.make_window("small", vio=False)
    g, o = solve_both(solver, win)
    assert_same_solution(g, o)
    assert np.all(g[1].dv == 0)


def test_no_fixed_frame_with_prior(solver):
    win = synth.make_window("small", n_fixed=0)
    assert_same_solution(*solve_both(solver, win))


def test_two_fixed_frames_fixed_cost(solver):
    """IMU/bias factors between two fixed keyframes and the prior on a fixed keyframe are fixed_cost, not cost."""
    win = synth.make_window("small", n_fixed=2)
    g, o = solve_both(solver, win)
    assert o[2]["fixed_cost"] > 0
    assert_same_solution(g, o)


def test_reference_inertial_optimisation(solver):  # imu_test.cpp:464-487
    win = rf.free_fall_window()
    g, o = solve_both(solver, win)
    rf.check_free_fall(win, g[1])
    assert_same_solution(g, o)


def test_reference_bias_estimation(solver):  # imu_test.cpp:545-568
    win = rf.bias_window()
    g, o = solve_both(solver, win)
    rf.check_bias(win, g[1])
    assert g[2]["termination"] == o[2]["termination"]


def test_reference_euroc_bias_run(solver):  # imu_test.cpp:885-945
    """29 consecutive solves on the growing EuRoC window through B200Optimizer (solve + write-back + biasDeltaCorrection): the
    reference's own assertion holds, and every keyframe ends where the oracle-driven run ends."""
    opt = api.B200Optimizer()
    kfs, stats = rf.euroc_bias_run(opt)
    rf.check_euroc_bias(kfs)
    kfs0, stats0 = rf.euroc_bias_run(rf.OracleOptimizer())
    assert [s["iterations"] for s in stats] == [s["iterations"] for s in stats0]
    assert [s["termination"] for s in stats] == [s["termination"] for s in stats0]
    for a, b in zip(kfs, kfs0):
        for name in ("ba", "bg", "v", "T_f_w", "dR", "dv", "dp"):
            x, y = orc.imu_get(a["imu"], name), orc.imu_get(b["imu"], name)
            assert np.abs(x - y).max() <= 1e-6 * max(np.abs(y).max(), 1e-3), name


@pytest.mark.parametrize("vio", [True, False])
def test_dense_marginal_prior(solver, vio):
    """a10: MarginalizationFactor with kept landmarks in the reduced system (…Analytic.cpp:341-383)."""
    win = synth.add_dense_prior(synth.make_window("small", vio=vio), n_keep=15, with_frame=vio)
    g, o = solve_both(solver, win)
    assert_same_solution(g, o)
    assert g[2]["n_reduced"] == o[2]["n_reduced"]


def test_c3_with_dense_prior_on_100_landmarks(solver):
    """SURVEY §8(d) C3 variant: dense marginal prior on (oldest free KF + 100 landmarks) -> n = 735 + ~260."""
    win = synth.add_dense_prior(synth.make_window("C3"), n_keep=100)
    g, o = solve_both(solver, win)
    assert_same_solution(g, o)


def test_sparsified_prior_vio(solver):
    """a11 VIO: IMUPriordx (incl. its un-whitened v/ba/bg Jacobians) + PoseToLandmark pseudo-observations."""
    win = synth.add_sparse_prior_vio(synth.make_window("small"), 15)
    g, o = solve_both(solver, win)
    assert_same_solution(g, o)
    win = synth.add_sparse_prior_vio(synth.make_window("C2", factor_kind=1), 60)
    assert_same_solution(*solve_both(solver, win))


def test_sparsified_prior_vo_chain(solver):
    """a11 VO: Landmark3DPrior + LandmarkToLandmark chain; chained landmarks are part of the reduced system."""
    win = synth.add_sparse_prior_vo(synth.make_window("small", vio=False), 12)
    g, o = solve_both(solver, win)
    assert_same_solution(g, o)
    assert g[2]["n_reduced"] == 6 * 9 + 3 * 12


def test_config4_bimono_nofov_sparsified(solver):
    """BASELINE config 4: 30 KF, non-overlapping cameras, localMapBA with the sparsified VO prior (99-link chain)."""
    win = synth.make_c4()
    assert win.n_frames == 30 and win.n_lmks == 4000 and win.n_obs == 24000 and not win.vio
    assert_same_solution(*solve_both(solver, win))


def test_ragged_and_degenerate_inputs(solver):
    win = synth.make_window("tiny")
    # a landmark nobody observes, a landmark seen only from the fixed keyframe, a single-observation landmark
    keep = np.ones(win.n_obs, bool)
    keep[win.obs_lmk == 3] = False
    idx = np.nonzero(win.obs_lmk == 5)[0]
    keep[idx[1:]] = False
    for name in ("obs_lmk", "obs_frame", "obs_cam", "obs_bearing", "obs_uv"):
        setattr(win, name, np.ascontiguousarray(getattr(win, name)[keep]))
    g, o = solve_both(solver, win)
    assert_same_solution(g, o)
    assert np.all(g[1].dlmk[3] == 0)


def test_interleaved_feature_order(solver):
    """lmk->getFeatures() order is arbitrary: here all left-camera features come before the right-camera ones, so a keyframe
    re-appears non-adjacently inside a landmark (general slot grouping path of the upload)."""
    win = synth.make_window("small")
    order = np.lexsort((np.arange(win.n_obs), win.obs_cam, win.obs_lmk))
    for name in ("obs_lmk", "obs_frame", "obs_cam", "obs_bearing", "obs_uv"):
        setattr(win, name, np.ascontiguousarray(getattr(win, name)[order]))
    f0 = win.obs_frame[win.obs_lmk == 0]
    assert len(f0) == 8 and f0[0] == f0[4] and f0[0] != f0[1]  # the keyframe re-appears after three others
    g, o = solve_both(solver, win)
    assert_same_solution(g, o)
    solver.upload(win)
    r, Jp, Jl, _ = solver.eval_visual(None)
    r0, Jp0, Jl0, _ = orc.eval_visual(win, None)
    assert rel(r, r0) < 1e-12 and rel(Jp, Jp0) < 1e-12


def test_invalid_inputs_are_rejected(solver):
    win = synth.make_window("tiny")
    bad = synth.make_window("tiny")
    bad.obs_lmk = bad.obs_lmk[::-1].copy()  # not landmark-major
    with pytest.raises(RuntimeError):
        solver.solve_window(bad)
    bad = synth.make_window("tiny")
    bad.obs_frame[0] = 99
    with pytest.raises(RuntimeError):
        solver.solve_window(bad)
    solver.solve_window(win)  # the handle stays usable


def test_visibility_indices_bit_exact(solver):
    """The flattening contract: (landmark, frame, camera) triplets reach the kernels unchanged."""
    win = synth.make_window("small")
    solver.upload(win)
    r, Jp, Jl, _ = solver.eval_visual(None)
    r0, Jp0, Jl0, _ = orc.eval_visual(win, None)
    # per-observation outputs are in the caller's observation order: any index permutation would break equality
    assert np.array_equal(np.argsort(np.abs(r[:, 0])), np.argsort(np.abs(r0[:, 0])))


def test_optimizer_interface_writeback(solver):
    """B200Optimizer mirrors AOptimizer::localMapVIOptimization incl. the state write-back (AOptimizer.cpp:391-434)."""
    win = synth.make_window("small")
    ref = synth.make_window("small")
    opt = api.B200Optimizer()
    assert opt.localMapVIOptimization(win, 1) is True
    rc0, d0, _ = orc.solve_window(ref)
    new = synth.apply_delta(ref, d0)
    assert np.abs(win.T_f_w - new["T_f_w"]).max() < 1e-9
    assert np.abs(win.ba - new["ba"]).max() < 1e-9
    # biasDeltaCorrection applied with the previous keyframe's delta (IMU.cpp:104-108)
    st = orc.imu_state(np.zeros(3), np.zeros(3))
    p = 0
    i = int(ref.imu_i[p])
    s = np.zeros(169)
    s[28:37], s[37:40], s[40:43] = ref.imu_dR[p], ref.imu_dv[p], ref.imu_dp[p]
    s[124:133], s[133:142], s[142:151], s[151:160], s[160:169] = (ref.imu_J_dR_bg[p], ref.imu_J_dv_ba[p], ref.imu_J_dv_bg[p],
                                                                  ref.imu_J_dp_ba[p], ref.imu_J_dp_bg[p])
    s2 = orc.bias_delta_correction(s, d0.dba[i], d0.dbg[i])
    assert np.abs(win.imu_dR[p] - s2[28:37]).max() < 1e-12
    assert np.abs(win.imu_dv[p] - s2[37:40]).max() < 1e-12
    assert np.abs(win.imu_dp[p] - s2[40:43]).max() < 1e-12


def test_writeback_across_a_long_gap(solver):
    """A keyframe more than 1 s after its previous keyframe: no IMU factor (AOptimizer.cpp:69) but its pre-integrated deltas
    are still corrected with the previous keyframe's dba / dbg (AOptimizer.cpp:421-434 has no dt test)."""
    F = synth.make_window("small").n_frames
    f_gap = F - 5
    win = synth.skip_imu_factor(synth.make_window("small"), f_gap)
    ref = synth.skip_imu_factor(synth.make_window("small"), f_gap)
    orig = synth.skip_imu_factor(synth.make_window("small"), f_gap)
    opt = api.B200Optimizer()
    assert opt.localMapVIOptimization(win, 1) is True
    rc0, d0, st0 = orc.solve_window(ref)
    assert rc0 == 0 and opt.last_stats["iterations"] == st0["iterations"]
    api.write_back(ref, d0, True)
    assert np.abs(win.T_f_w - ref.T_f_w).max() < 1e-9 and np.abs(win.bg - ref.bg).max() < 1e-9
    sk, sk0, sko = win.skipped_preint, ref.skipped_preint, orig.skipped_preint
    i = int(sko.prev[0])
    assert int(sko.frame[0]) == f_gap and np.abs(d0.dbg[i]).max() > 0
    dp = sko.dp[0] + sko.J_dp_ba[0].reshape(3, 3) @ d0.dba[i] + sko.J_dp_bg[0].reshape(3, 3) @ d0.dbg[i]
    dR = sko.dR[0].reshape(3, 3) @ synth.exp_so3(sko.J_dR_bg[0].reshape(3, 3) @ d0.dbg[i])
    assert np.abs(sk0.dp[0] - dp).max() < 1e-15 and np.abs(sk0.dR[0] - dR.reshape(9)).max() < 1e-15
    assert np.abs(sk.dp[0] - dp).max() < 1e-9 and np.abs(sk.dR[0] - dR.reshape(9)).max() < 1e-9 and np.abs(sk.dv[0] - sk0.dv[0]).max() < 1e-9
    assert np.abs(sk.dp[0] - sko.dp[0]).max() > 1e-12


# ---------------------------------------------------------------------------------------------------------------------
# the factorisation / assembly variants must agree with each other and with the oracle
# ---------------------------------------------------------------------------------------------------------------------
def _solve_with_env(win, env, cfg=None):
    import os
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        s = api.Solver(cfg) if cfg is not None else api.Solver()
        out = s.solve_window(win)   # the environment is read at upload time
        s.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return out


@pytest.mark.parametrize("name", ["small", "C2", "C3"])
def test_band_cholesky_matches_cluster_cholesky(name):
    """k_chol_band (single CTA, banded; the default when the band fits) against the 16-CTA cluster factorisation."""
    win = synth.make_window(name)
    rc_b, d_b, st_b = _solve_with_env(win, {})
    rc_c, d_c, st_c = _solve_with_env(win, {"SDV_CHOL_VARIANT": "4"})
    assert rc_b == rc_c == 0
    assert st_b["iterations"] == st_c["iterations"] and st_b["termination"] == st_c["termination"]
    assert st_b["trace_accepted"] == st_c["trace_accepted"]
    for a, b in ((d_b.dpose, d_c.dpose), (d_b.dv, d_c.dv), (d_b.dba, d_c.dba), (d_b.dbg, d_c.dbg), (d_b.dlmk, d_c.dlmk)):
        assert np.abs(a - b).max() <= 1e-9 * np.abs(b).max() + 1e-13


def test_wide_band_falls_back_to_cluster_cholesky(solver):
    """Landmarks seen from 10 keyframes: half-bandwidth > 7 blocks of 16, the banded kernel does not apply."""
    win = synth.make_window("C2", span=10)
    g, o = solve_both(solver, win)
    assert_same_solution(g, o)
    assert_same_states(win, g[1], o[1])


def test_dense_small_window_band_kernel(solver):
    """A 6-keyframe window whose reduced system is dense inside the band (every landmark seen from 3 of 6 frames)."""
    win = synth.make_window("tiny")
    g, o = solve_both(solver, win)
    assert_same_solution(g, o)


@pytest.mark.parametrize("name", ["small", "C2", "C3"])
def test_fused_schur_does_not_depend_on_the_tiling(name):
    """k_lin_schur sums the landmarks of a tile in shared memory before its atomics reach S, and finds the runs of landmarks
    seen from the same keyframes inside each tile: tiles of at most 8 slots (one or two landmarks, hardly any run) against the
    default tiling."""
    win = synth.make_window(name)
    rc_a, d_a, st_a = _solve_with_env(win, {})
    rc_b, d_b, st_b = _solve_with_env(win, {"SDV_FUSED_TILE_SLOTS": "8"})
    assert rc_a == rc_b == 0 and st_a["iterations"] == st_b["iterations"]
    for a, b in ((d_a.dpose, d_b.dpose), (d_a.dv, d_b.dv), (d_a.dlmk, d_b.dlmk)):
        assert np.abs(a - b).max() <= 1e-9 * np.abs(b).max() + 1e-13
    rc0, d0, st0 = orc.solve_window(win, mode=0, nthreads=8)
    assert np.abs(d_a.dpose - d0.dpose).max() <= 1e-6 * np.abs(d0.dpose).max()


def test_gradient_tolerance_termination_matches_oracle(solver):
    """The gradient test lives in the prologue of k_chol_band (and in k_sysprep for the other variants)."""
    win = synth.make_window("small")
    cfg = api.default_config()
    cfg.gradient_tolerance = 1e3   # met at iteration 0 after the first linearisation
    g, o = solve_both(solver, win, cfg)
    assert g[2]["termination"] == o[2]["termination"]
    assert g[2]["iterations"] == o[2]["iterations"]


def test_band_of_one_block(solver):
    """BA-only window with two-keyframe tracks: consecutive 6-dof poses couple, half-bandwidth of ONE 16-column block —
    the streaming-update warps of k_chol_band have no (2,1) task to wait for."""
    win = synth.make_window("C2", span=2, vio=False)
    g, o = solve_both(solver, win)
    assert_same_solution(g, o)
    assert_same_states(win, g[1], o[1])


@pytest.mark.parametrize("name,cfg,kind", [("tiny_vio_angular", "tiny", 0), ("tiny_vio_pixel", "tiny", 1), ("small_vio_angular", "small", 0)])
def test_golden_solutions(solver, name, cfg, kind):
    """CUDA path against the committed golden solutions (tests/golden/*.npz, generated by tests/golden/make_golden.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    win = synth.make_window(cfg, factor_kind=kind)
    assert np.array_equal(win.obs_lmk, g["obs_lmk"]) and np.array_equal(win.obs_frame, g["obs_frame"])  # visibility bit-exact
    rc, d, st = solver.solve_window(win)
    assert rc == 0 and st["iterations"] == int(g["iterations"]) and st["termination"] == str(g["termination"])
    assert list(st["trace_accepted"]) == list(g["trace_accepted"])
    for k in ("dpose", "dv", "dba", "dbg"):
        a, b = getattr(d, k), g[k]
        assert np.abs(a - b).max() <= 1e-6 * np.abs(b).max() + 1e-13


def test_full_solve_c5_band_kernel(solver):
    """BASELINE.json config 5 on one GPU: 200 KF x 100k landmarks x 800k observations; the reduced system (n = 2985, 187
    blocks of 16, half-bandwidth 4) is factored by the banded single-CTA kernel."""
    win = synth.make_window("C5")
    assert (win.n_frames, win.n_lmks, win.n_obs) == (200, 100000, 800000)
    g, o = solve_both(solver, win)
    assert_same_solution(g, o)
    assert_same_states(win, g[1], o[1])


def test_next_window_with_marginal_prior_from_the_oracle(solver):
    """The prior the restated marginalisation driver (oracle/marginalize.py, row a15 groundwork) produces for the window without
    its oldest keyframe — 15 + 3 x 43 columns, rank-deficient by the 6-dof gauge — through the CUDA dense-prior path (a10)."""
    from oracle import marginalize

    win = synth.make_window("small")
    prior, info = marginalize.marginalize_oldest(win)
    w2 = marginalize.drop_oldest_frame(win, prior)
    assert w2.dense_prior is not None and w2.dense_prior.J.shape == (info["n_full"], 15 + 3 * len(info["keep"]))
    g, o = solve_both(solver, w2)
    assert_same_solution(g, o)
    assert_same_states(w2, g[1], o[1])


def test_next_window_with_sparsified_prior_from_the_oracle(solver):
    """sparsifyVIO restated by the oracle (oracle/marginalize.py): IMUPriordx on the oldest remaining keyframe + 43
    PoseToLandmark factors with the information the marginal covariance gives them, through the CUDA sparse-prior path (a11)."""
    from oracle import marginalize

    win = synth.make_window("small")
    prior, info = marginalize.marginalize_oldest(win)
    w2 = marginalize.with_sparse_prior(win, marginalize.sparsify_vio(win, info))
    g, o = solve_both(solver, w2)
    assert_same_solution(g, o)
    assert_same_states(w2, g[1], o[1])


def test_next_vo_window_with_sparsified_prior_from_the_oracle(solver):
    """sparsifyVO restated by the oracle: Landmark3DPrior on the landmark of least entropy + the LandmarkToLandmark chain, through
    the CUDA sparse-prior path (a11), on a window without IMU and without a fixed keyframe (the prior alone holds the gauge)."""
    from oracle import marginalize

    win = synth.make_window("small", vio=False)
    prior, info = marginalize.marginalize_oldest(win)
    w2 = marginalize.with_sparse_prior(win, marginalize.sparsify_vo(win, info))
    g, o = solve_both(solver, w2)
    assert_same_solution(g, o)
    assert_same_states(w2, g[1], o[1])


# ---------------------------------------------------------------------------------------------------------------------
# cases queued at the end of round 1 (VERDICT r01 "untested on GPU")
# ---------------------------------------------------------------------------------------------------------------------
def test_c1_euroc_window(solver):
    """BASELINE config 1 on the GPU: 10 keyframes 0.5 s apart on the EuRoC ground truth, stereo rig of eth.yaml, 10 k obs."""
    win = synth.make_window("C1")
    g, o = solve_both(solver, win)
    assert_same_solution(g, o)
    assert_same_states(win, g[1], o[1])
    assert g[2]["final_cost"] < 0.01 * g[2]["initial_cost"]
    win = synth.make_window("C1", factor_kind=1)
    g, o = solve_both(solver, win)
    assert_same_solution(g, o)


def test_keyframe_without_imu_in_a_vio_window(solver):
    """getIMU() == nullptr on one keyframe of a VIO window: no v / ba / bg blocks, no IMU factor on either side of it
    (AOptimizer.cpp:30-52, 60-73); its v / ba / bg updates stay exactly zero."""
    win = synth.make_window("small")
    f = 4
    keep = (win.imu_i != f) & (win.imu_j != f)
    for k in synth.IMU_FACTOR_FIELDS:
        setattr(win, k, getattr(win, k)[keep].copy())
    win.has_imu[f] = 0
    win.normalise()
    g, o = solve_both(solver, win)
    assert_same_solution(g, o)
    d = g[1]
    assert np.all(d.dv[f] == 0) and np.all(d.dba[f] == 0) and np.all(d.dbg[f] == 0) and np.abs(d.dpose[f]).max() > 0


def test_chained_dense_priors_with_resurrected_landmarks(solver):
    """The back end's loop over several keyframes — solve, write back, marginalise the oldest keyframe (the previous prior is
    folded in, …Analytic.cpp:631-660), drop it — with the window solves on the GPU and the marginalisation by the oracle.
    Landmarks kept by a prior while no remaining keyframe sees them ("supposed to be in the map", …Analytic.cpp:369-373) have a
    parameter block and no observation.  Every solve must match the oracle solve of the same window."""
    from oracle import marginalize

    w = synth.make_window("small")
    n_prior_only = 0
    for step in range(5):
        prior, info = marginalize.marginalize_oldest(w)
        assert prior is not None
        w = marginalize.drop_oldest_frame(w, prior)
        if step == 0:   # three kept landmarks lose their tracks: from here on only the prior knows them
            lost = np.isin(w.obs_lmk, w.dense_prior.keep_lmk[[1, 5, 9]])
            for name in ("obs_lmk", "obs_frame", "obs_cam", "obs_bearing", "obs_uv"):
                setattr(w, name, np.ascontiguousarray(getattr(w, name)[~lost]))
        seen = np.zeros(w.n_lmks, bool)
        seen[w.obs_lmk] = True
        n_prior_only += int((~seen[w.dense_prior.keep_lmk]).sum())
        g, o = solve_both(solver, w)
        assert_same_solution(g, o)
        assert_same_states(w, g[1], o[1])
        api.write_back(w, g[1], True)
    assert n_prior_only > 0


def _trim_landmarks(win, n_drop):
    """The same window with its last `n_drop` landmarks (and their observations) gone: what a back end sees from one keyframe
    to the next — same keyframes, a few landmarks more or less."""
    L = win.n_lmks - n_drop
    keep = win.obs_lmk < L
    for name in ("obs_lmk", "obs_frame", "obs_cam", "obs_bearing", "obs_uv"):
        setattr(win, name, np.ascontiguousarray(getattr(win, name)[keep]))
    win.lmk_t = np.ascontiguousarray(win.lmk_t[:L])
    return win


def test_one_cuda_graph_serves_consecutive_windows():
    """The whole solve is one CUDA graph whose kernel arguments do not depend on the window (the problem description is read
    from device memory, scratch is laid out from capacities, grids are persistent upper bounds): windows of different
    landmark / observation counts replay the graph captured for the first one, and each still matches the oracle.  (One graph
    per number of LM iterations per trip of its WHILE node — chosen from the previous solve's iteration count, at most 4 — is
    kept: the first window may add one, later windows add none.)"""
    s = api.Solver()
    drops = [0, 37, 5, 120, 64, 0]
    iters = set()
    for n_drop in drops:
        win = _trim_landmarks(synth.make_window("C2"), n_drop)
        g = s.solve_window(win)
        o = orc.solve_window(win, mode=0, nthreads=8)
        assert_same_solution(g, o)
        iters.add(g[2]["iterations"])
        if n_drop == 37:
            builds = s.graph_builds()
    assert builds <= 2 and (len(iters) > 1 or s.graph_builds() == builds), (builds, s.graph_builds(), iters)
    assert s.graph_builds() <= 1 + len(iters)
    # a different keyframe count is a different graph
    b0 = s.graph_builds()
    s.solve_window(synth.make_window("small"))
    assert s.graph_builds() == b0 + 1
    s.close()


def test_alternating_kinds_of_solves_replay_parked_graphs():
    """The front-end optimizer instance runs landmarkOptimization and the single-frame solves in turn on ONE handle: every kind has
    its own launch signature, and the graphs that fall out of use are parked instead of destroyed — after the first round no solve
    captures or instantiates anything, and every solve still matches the oracle."""
    s = api.Solver()
    base = synth.make_window("small")
    kinds = [lambda: api.landmark_window(synth.make_window("small"), 3)[0], lambda: api.single_frame_window(synth.make_window("small"), 0, False)[0],
             lambda: api.single_frame_window(synth.make_window("small"), 0, True)[0], lambda: synth.make_window("small")]
    builds = []
    for rnd in range(3):
        for mk in kinds:
            win = mk()
            g = s.solve_window(win)
            o = orc.solve_window(win, mode=0, nthreads=4)
            assert g[2]["iterations"] == o[2]["iterations"] and g[2]["termination"] == o[2]["termination"]
            assert np.abs(g[1].dpose - o[1].dpose).max() <= 1e-6 * max(np.abs(o[1].dpose).max(), 1e-12)
        builds.append(s.graph_builds())
    assert builds[2] == builds[1], builds          # (round 1 may still add the graphs of the iteration counts it learned in round 0)
    assert builds[1] <= 2 * len(kinds), builds
    del base
    s.close()


def test_parallel_host_structure_pass_matches_the_serial_one():
    """Windows with >= 16384 observations build their slot lists on four host threads (observation list cut at landmark
    boundaries); SDV_NO_HOST_POOL=1 forces the serial pass.  Same LM trace and solution, also when a keyframe re-appears
    non-adjacently inside a landmark (general grouping path) and when some landmarks have no observation at all."""
    win = synth.make_window("C3")
    order = np.lexsort((np.arange(win.n_obs), win.obs_cam, win.obs_lmk))      # left-camera features first inside every landmark
    keep = np.ones(win.n_obs, bool)
    keep[np.isin(win.obs_lmk, [0, 1, 4999, 5000, 9999])] = False              # empty landmarks, also at the part boundaries
    order = order[keep[order]]
    for name in ("obs_lmk", "obs_frame", "obs_cam", "obs_bearing", "obs_uv"):
        setattr(win, name, np.ascontiguousarray(getattr(win, name)[order]))
    a = _solve_with_env(win, {})
    b = _solve_with_env(win, {"SDV_NO_HOST_POOL": "1"})
    assert_same_solution(a, b, tol=1e-9)
    o = orc.solve_window(win, mode=0, nthreads=8)
    assert_same_solution(a, o)
    # invalid input is still rejected by the parallel pass
    bad = synth.make_window("C3")
    bad.obs_frame[41234] = bad.n_frames
    with pytest.raises(RuntimeError):
        api.Solver().solve_window(bad)
    bad = synth.make_window("C3")
    bad.obs_lmk[60000] = 3                                                     # not landmark-major
    with pytest.raises(RuntimeError):
        api.Solver().solve_window(bad)


# ---------------------------------------------------------------------------------------------------------------------
# IMU pre-integration on the device: sdv_preintegrate against the oracle's processIMU chain
# ---------------------------------------------------------------------------------------------------------------------
def _oracle_preint_chain(T_kf, v_kf, ba, bg, acc, gyr, dts, eta, rate, dR_stale=None):
    kf = orc.imu_state(acc[0], gyr[0], T_f_w=T_kf, v=v_kf, ba=ba, bg=bg, is_kf=True)
    if dR_stale is not None:
        kf[28:37] = dR_stale
    last = kf
    for k in range(len(dts)):
        nxt_acc, nxt_gyr = (acc[k + 1], gyr[k + 1]) if k + 1 < len(dts) else (acc[k], gyr[k])
        last = orc.process_imu(last, ba, bg, dts[k], eta, rate, nxt_acc, nxt_gyr)
    return last


def test_preintegration_kernel_matches_the_oracle(solver):
    rng = np.random.default_rng(5)
    n, rate = 49, 200.0
    counts = rng.integers(30, 70, n)
    counts[3], counts[7] = 1, 0                                     # a one-sample interval and an empty one
    sp = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    S = int(sp[-1])
    acc = rng.normal(0, 2.0, (S, 3)) + [0, 0, 9.81]
    gyr = rng.normal(0, 0.3, (S, 3))
    gyr[sp[5]:sp[6]] = 0.0                                          # zero rotation: the small-angle branches of exp / Jr
    dts = np.full(S, 1.0 / rate)
    dts[sp[10] + 4] = 1.7                                           # > 1 s: replaced by 1 / rate (IMU.cpp:23-25)
    T = np.stack([synth.T34(np.vstack([np.hstack([synth.exp_so3(rng.normal(0, 0.5, 3)), rng.normal(0, 2, (3, 1))]), [0, 0, 0, 1]])) for _ in range(n)])
    v, ba, bg = rng.normal(0, 1, (n, 3)), rng.normal(0, 0.05, (n, 3)), rng.normal(0, 0.01, (n, 3))
    stale = np.stack([synth.exp_so3(rng.normal(0, 0.2, 3)).reshape(9) for _ in range(n)])
    eta = rf.eta(rate)
    out = solver.preintegrate(sp, acc, gyr, dts, T, v, ba, bg, eta, rate, dR_stale=stale)
    names = (("dR", "dR"), ("dv", "dv"), ("dp", "dp"), ("cov", "Sigma"), ("J_dR_bg", "J_dR_bg"), ("J_dv_ba", "J_dv_ba"), ("J_dv_bg", "J_dv_bg"),
             ("J_dp_ba", "J_dp_ba"), ("J_dp_bg", "J_dp_bg"), ("T_pred", "T_f_w"), ("v_pred", "v"))
    for k in range(n):
        if counts[k] == 0:
            assert np.all(out["cov"][k] == 0) and np.allclose(out["dR"][k], np.eye(3).reshape(9)) and np.all(out["dp"][k] == 0)
            continue
        a, b = sp[k], sp[k + 1]
        ref = _oracle_preint_chain(T[k], v[k], ba[k], bg[k], acc[a:b], gyr[a:b], dts[a:b], eta, rate, stale[k])
        for mine, theirs in names:
            r = orc.imu_get(ref, theirs)
            assert np.abs(out[mine][k] - r).max() <= 1e-11 * max(1.0, np.abs(r).max()), (k, mine)
    # the reference's own known-answer test, imu_test.cpp:948-995 (two half-second steps)
    acc2, gyr2 = np.array([[0.1, 9.81, 0.0]] * 2), np.array([[0.0, 0.0, 0.1]] * 2)
    o2 = solver.preintegrate([0, 2], acc2, gyr2, [0.5, 0.5], np.eye(3, 4).reshape(1, 12), np.zeros((1, 3)), np.zeros((1, 3)), np.zeros((1, 3)), rf.eta(200), 200)
    imu0 = orc.imu_state(acc2[0], gyr2[0], is_kf=True)
    imu1 = orc.process_imu(imu0, np.zeros(3), np.zeros(3), 0.5, rf.eta(200), 200, acc2[1], gyr2[1])
    imu2 = orc.process_imu(imu1, np.zeros(3), np.zeros(3), 0.5, rf.eta(200), 200, acc2[1], gyr2[1])
    assert np.abs(o2["dp"][0] - orc.imu_get(imu2, "dp")).max() < 1e-13
    assert np.abs(o2["cov"][0] - orc.imu_get(imu2, "Sigma")).max() <= 1e-11 * np.abs(orc.imu_get(imu2, "Sigma")).max()

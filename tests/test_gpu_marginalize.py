"""Marginal-prior construction on the GPU: sdv_marginalize / sdv_schur_prior against the
oracle's restatement (oracle/marginalize.py, oracle/marg.hpp) and against numpy.

What can be compared entry by entry is: the information matrix A and gradient b, the Schur complement Ak / bk, the
spectrum.  Eigenvectors are only defined up to sign (and up to a rotation inside a degenerate eigenspace), so J and r0 are
compared through the invariants J^T J = Ak, J^T r0 = -bk — and through what they are FOR: the next window solved with the
GPU-made prior must land where it lands with the oracle-made prior (<= 1e-6 relative)."""
import numpy as np
import pytest

from oracle import marginalize, oracle
from sadvio_b200 import abi, synth
from tests import test_oracle_marginalization as toy

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(1e-300, np.abs(np.asarray(b)).max()))


# ------------------------------------------------------------------------------------------------ the dense core alone
@pytest.mark.parametrize("n,rank", [(7, 7), (40, 40), (60, 23), (150, 150), (333, 333), (420, 200)])
def test_block_jacobi_eigensolver_against_numpy(solver, n, rank):
    """sdv_schur_prior with m = 0 is the rank-revealing decomposition alone: Lambda / U of a symmetric PSD matrix."""
    rng = np.random.default_rng(n * 100 + rank)
    B = rng.normal(size=(n, rank))
    A = B @ B.T + (0.0 if rank < n else 1e-3) * np.eye(n)
    out = solver.schur_prior(A, np.zeros(n), m=0)
    ref = np.linalg.eigvalsh(A)
    scale = np.abs(ref).max()
    big = ref[ref > 1e-9 * scale]
    assert out["n_full"] >= big.size and np.all(np.diff(out["Lambda"]) >= 0)
    assert np.abs(out["Lambda"][-big.size:] - big).max() <= 1e-12 * scale * n
    U = out["U"][:, -big.size:]
    assert np.abs(U.T @ U - np.eye(big.size)).max() <= 1e-11 * n
    assert np.abs((U * out["Lambda"][-big.size:]) @ U.T - A).max() <= 1e-11 * scale * n
    assert 0 < out["eig_sweeps_n"] < 25
    print(f"n = {n}, rank {rank}: {out['eig_sweeps_n']} sweeps")


def test_reference_toy_graph_through_the_abi(solver):
    """cpp/tests/marginalization_test.cpp:219-223, :300-313 — _m = 9, _n = 6, |Ak| = 36, symmetric, coupled landmarks."""
    A, b = toy.information()
    out = solver.schur_prior(A, b, m=9)
    assert out is not None and out["m"] == 9 and out["n"] == 6
    Ak = out["Ak"]
    assert Ak.size == 36 and np.linalg.norm(Ak - Ak.T) < 1e-8 and abs(np.trace(Ak[0:3, 3:6])) > 0
    assert np.allclose(out["bk"], 0)
    assert solver.schur_prior(A[:12, :12], b[:12], m=9) is None          # n = 3 < 4 (:215, test :334)


@pytest.mark.parametrize("noise", [0.0, 0.7])
def test_schur_complement_invariants_on_the_toy_graph(solver, noise):
    A, b = toy.information(noise=noise, seed=3, lmks=toy.LMKS_GENERIC)
    out = solver.schur_prior(A, b, m=9)
    Ak_np, bk_np = toy.numpy_reference(A, b, 9)
    scale = np.abs(Ak_np).max()
    assert np.abs(out["Ak"] - Ak_np).max() <= 1e-9 * scale
    assert np.abs(out["bk"] - bk_np).max() <= 1e-9 * max(1.0, np.abs(bk_np).max())
    lam = np.linalg.eigvalsh(Ak_np)
    big = lam[lam > 1e-6 * scale]
    assert big.size == 3 and out["n_full"] >= 3
    assert np.all(out["Lambda"] > 1e-12) and np.allclose(out["Lambda"][-3:], big, rtol=1e-9) and np.all(out["Lambda"][:-3] < 1e-8 * scale)
    J, r0 = out["J"], out["r0"]
    assert np.abs(J.T @ J - Ak_np).max() <= 1e-8 * scale
    assert np.abs(J.T @ r0 + bk_np).max() <= 1e-8 * max(1.0, np.abs(bk_np).max())
    # a marginalised parameter without information is ignored by the thresholded inverse (marginalization.cpp:234-240)
    A2, b2 = np.zeros((A.shape[0] + 1,) * 2), np.zeros(A.shape[0] + 1)
    A2[1:, 1:], b2[1:] = A, b
    o2 = solver.schur_prior(A2, b2, m=10)
    assert np.abs(out["Ak"] - o2["Ak"]).max() <= 1e-9 * scale


# ------------------------------------------------------------------------------------------------ the whole marginalize()
def _compare_with_oracle(solver, win, sparsify=False):
    dense, sparse, info = solver.marginalize(win, sparsify)
    prior0, info0 = marginalize.marginalize_oldest(win)
    assert (dense is None) == (prior0 is None)
    assert info["m"] == info0["m"] and info["n"] == info0["n"]
    assert list(info["keep"]) == list(info0["keep"]) and list(info["marg"]) == list(info0["marg"])
    scaleA = np.abs(info0["A"]).max()
    assert np.abs(info["A"] - info0["A"]).max() <= 1e-12 * scaleA, "information matrix"
    assert np.abs(info["A"] - info["A"].T).max() <= 1e-13 * scaleA
    assert np.abs(info["b"] - info0["b"]).max() <= 1e-11 * max(1.0, np.abs(info0["b"]).max()), "gradient"
    scale = np.abs(info0["Ak"]).max()
    # Arm Amm^+ Arm^T carries the conditioning of the marginalised block (a chained prior leaves weakly constrained velocity /
    # bias directions in it): the forward error of ANY solver is ~ cond(Amm) eps
    m = info0["m"]
    wm = np.linalg.eigvalsh(0.5 * (info0["A"][:m, :m] + info0["A"][:m, :m].T))
    cond = wm[-1] / wm[wm > 1e-12].min()
    tol_ak = max(1e-9, 100 * 2.2e-16 * cond)
    assert np.abs(info["Ak"] - info0["Ak"]).max() <= tol_ak * scale, ("Schur complement", cond)
    assert np.abs(info["bk"] - info0["bk"]).max() <= tol_ak * max(1.0, np.abs(info0["bk"]).max())
    # the significant part of the spectrum, and the prior's invariants
    lam0 = np.asarray(info0["Lambda"])
    nbig = int((lam0 > 1e-7 * scale).sum())
    # (a symmetric eigen-solver is backward stable: eigenvalues carry an ABSOLUTE error of a few eps |Ak|, whichever solver)
    assert np.abs(info["Lambda"][-nbig:] - lam0[-nbig:]).max() <= max(1e-12 * info["n"], 2 * tol_ak) * scale
    assert np.abs(dense.J.T @ dense.J - info["Ak"]).max() <= 1e-8 * scale
    assert np.abs(dense.J.T @ dense.r0 + info["bk"]).max() <= 1e-8 * max(1.0, np.abs(info["bk"]).max())
    assert np.abs(dense.J.T @ dense.J - info0["Ak"]).max() <= max(1e-8, tol_ak) * scale
    assert dense.frame == prior0.frame and np.array_equal(dense.keep_lmk, prior0.keep_lmk) and np.array_equal(dense.keep_col, prior0.keep_col)
    return dense, sparse, info, prior0, info0


def _next_window_parity(solver, win, prior_gpu, prior_orc, tol=5e-6):
    """The shorter window solved (on the GPU) with either prior: same LM trace, same states.  The two priors agree on
    everything that carries information (J^T J, J^T r0: asserted above); they differ in WHICH rounding-noise eigenpairs of a
    rank-deficient Ak came out above the reference's absolute 1e-12 threshold (the oracle's QL and the GPU's Jacobi disagree
    there, as two Eigen versions would).  Those rows add ~1e-9-scale information but (u^T bk)^2 / lambda to the cost the LM
    accept test sees, so two function-tolerance-terminated runs end a few 1e-6 apart; with a full-rank Ak (steady state) the
    states agree to 1e-9."""
    res = []
    for pr in (prior_gpu, prior_orc):
        w2 = marginalize.drop_oldest_frame(win, pr)
        rc, d, st = solver.solve_window(w2)
        assert rc == 0
        res.append((d, st))
    (d, st), (d0, st0) = res
    assert st["iterations"] == st0["iterations"] and st["termination"] == st0["termination"]
    for a, b in ((d.dpose, d0.dpose), (d.dv, d0.dv), (d.dba, d0.dba), (d.dbg, d0.dbg)):
        if np.abs(b).max() > 0:
            assert _rel(a, b) <= tol, _rel(a, b)
    assert abs(st["final_cost"] - st0["final_cost"]) <= tol * st0["final_cost"]


@pytest.mark.parametrize("name,kind,vio", [("small", 0, True), ("small", 1, True), ("small", 0, False), ("C2", 0, True)])
def test_marginalize_matches_oracle(solver, name, kind, vio):
    win = synth.make_window(name, factor_kind=kind, vio=vio)
    dense, _, info, prior0, info0 = _compare_with_oracle(solver, win)
    assert info["eig_sweeps_m"] > 0 and info["eig_sweeps_n"] > 0 and info["ms_device"] > 0
    _next_window_parity(solver, win, dense, prior0)


def test_pose_prior_on_frame_one_enters_for_the_angular_optimizer_only(solver):
    for kind in (0, 1):
        win = synth.make_window("small", factor_kind=kind)
        f1 = win.n_frames - 2
        win.has_prior[f1] = 1
        win.T_prior[f1] = win.T_f_w[f1] + 0.01
        win.inf_prior[f1] = 50.0
        _compare_with_oracle(solver, win)                               # (…Analytic.cpp:676-687 vs BundleAdjustment…:606-617)


def test_chained_marginalisation_with_resurrected_landmarks(solver):
    """Three consecutive keyframes leave the window: every marginalisation folds the previous prior in as one more block
    (…Analytic.cpp:631-660); kept landmarks nobody observes any more come back through it (marginalization.cpp:118-139)."""
    win = synth.make_window("C2")
    win_o = synth.make_window("C2")
    for step in range(3):
        dense, _, info, prior0, info0 = _compare_with_oracle(solver, win)
        if step > 0:
            assert len(info["keep"]) > 0
        # both chains continue with their OWN prior (the comparison above is per step, on identical inputs)
        assert np.array_equal(win.lmk_t, win_o.lmk_t)
        nxt = marginalize.drop_oldest_frame(win, prior0)
        # chained: Amm now holds the weakly constrained velocity / bias directions the previous prior left (cond ~ 1e9, printed
        # by _compare_with_oracle's tolerance): Arm Amm^+ Arm^T is uncertain at ~cond eps |Ak| for ANY solver, which is as large as
        # the smallest informative eigenvalues of Ak — the weak directions of the next solve move by ~1e-5 relative
        _next_window_parity(solver, win, dense, prior0, tol=2e-4)
        win = nxt
        win.n_fixed = 0
        win_o = marginalize.drop_oldest_frame(win_o, prior0)
        win_o.n_fixed = 0


def _full_rank_vio_window():
    """A VIO window whose oldest keyframe carries a full 15-dof dense prior (the steady state of a running back end: the
    previous marginalisation).  Without it Ak has a null space — frame 1 is tied to frame 0 by 15 IMU rows that also carry
    frame 0's nine velocity / bias unknowns — the reference keeps whichever rounding-noise eigenvalues come out above 1e-12,
    and their 1 / lambda dominates the sparsified covariance: solver noise, not comparable between two eigen-solvers."""
    win = synth.make_window("small", n_fixed=0)
    return synth.add_dense_prior(win, n_keep=20).normalise()


def test_sparsify_vio_matches_oracle(solver):
    win = _full_rank_vio_window()
    dense, sparse, info, prior0, info0 = _compare_with_oracle(solver, win, sparsify=True)
    assert info["n_full"] == info["n"] == info0["n_full"]                 # full rank: nothing is thresholded away
    assert _rel(info["Lambda"], info0["Lambda"]) <= 1e-9
    sp0 = marginalize.sparsify_vio(win, info0)
    assert sparse.has_imu_prior and sparse.frame == sp0.frame and np.array_equal(sparse.p2l_lmk, sp0.p2l_lmk)
    assert np.abs(sparse.p2l_delta - sp0.p2l_delta).max() <= 1e-12
    assert _rel(sparse.p2l_sqrt_inf, sp0.p2l_sqrt_inf) <= 1e-6
    assert _rel(sparse.imu_sqrt_inf, sp0.imu_sqrt_inf) <= 1e-6
    # and the next window solved with it
    res = []
    for sp in (sparse, sp0):
        w2 = marginalize.with_sparse_prior(win, sp)
        rc, d, st = solver.solve_window(w2)
        assert rc == 0
        res.append((d, st))
    assert res[0][1]["iterations"] == res[1][1]["iterations"] and _rel(res[0][0].dpose, res[1][0].dpose) <= 1e-6


def test_sparsify_vio_on_a_rank_deficient_marginal(solver):
    """The first marginalisation of a run (no previous prior): the kernels are checked on IDENTICAL inputs — the oracle's
    sparsifyVIO fed with the U / Lambda the GPU decomposition produced; the 15 x 15 frame factor inverts a numerically
    singular matrix there (see _full_rank_vio_window) and is only required to be symmetric and finite."""
    win = synth.make_window("small")
    dense, sparse, info, prior0, info0 = _compare_with_oracle(solver, win, sparsify=True)
    sp0 = marginalize.sparsify_vio(win, dict(n=info["n"], U=info["U"], Lambda=info["Lambda"], keep=list(info["keep"])))
    assert np.abs(sparse.p2l_delta - sp0.p2l_delta).max() <= 1e-12 and _rel(sparse.p2l_sqrt_inf, sp0.p2l_sqrt_inf) <= 1e-6
    S = sparse.imu_sqrt_inf.reshape(15, 15)
    assert np.all(np.isfinite(S)) and np.abs(S - S.T).max() <= 1e-9 * np.abs(S).max()
    rc, d, st = solver.solve_window(marginalize.with_sparse_prior(win, sparse))
    assert rc == 0 and st["final_cost"] <= st["initial_cost"]


def test_sparsify_vo_matches_oracle(solver):
    win = synth.make_window("small", vio=False)
    dense, sparse, info, prior0, info0 = _compare_with_oracle(solver, win, sparsify=True)
    sp0 = marginalize.sparsify_vo(win, info0)                             # (frame 0's pose prior fixes the gauge: Ak has full rank)
    assert sparse.has_lmk_prior and sparse.lmk0 == sp0.lmk0
    assert np.array_equal(sparse.l2l_a, sp0.l2l_a) and np.array_equal(sparse.l2l_b, sp0.l2l_b)
    assert np.abs(sparse.l2l_delta - sp0.l2l_delta).max() <= 1e-12 and np.allclose(sparse.lmk_prior, sp0.lmk_prior)
    assert _rel(sparse.l2l_sqrt_inf, sp0.l2l_sqrt_inf) <= 1e-6 and _rel(sparse.lmk_sqrt_inf, sp0.lmk_sqrt_inf) <= 1e-6
    w2 = marginalize.with_sparse_prior(win, sparse)
    w0 = marginalize.with_sparse_prior(win, sp0)
    (rc, d, st), (rc0, d0, st0) = solver.solve_window(w2), solver.solve_window(w0)
    assert rc == rc0 == 0 and st["iterations"] == st0["iterations"] and _rel(d.dpose, d0.dpose) <= 1e-6


def test_too_few_kept_parameters_returns_no_prior(solver):
    """A VO window whose oldest keyframe shares no landmark with the others: n = 0 < 4, marginalize() returns false (:215)."""
    win = synth.make_window("tiny", vio=False)
    f0 = win.n_frames - 1
    lm0 = np.unique(win.obs_lmk[win.obs_frame == f0])
    keep = ~(np.isin(win.obs_lmk, lm0) & (win.obs_frame != f0))          # frame 0's landmarks are seen by frame 0 only
    for k in ("obs_lmk", "obs_frame", "obs_cam", "obs_bearing", "obs_uv"):
        setattr(win, k, getattr(win, k)[keep].copy())
    dense, sparse, info = solver.marginalize(win.normalise())
    prior0, _ = marginalize.marginalize_oldest(win)
    assert dense is None and prior0 is None and info["ok"] == 0 and info["n"] < 4


def test_c3_marginalisation_timing_is_reported(solver):
    """The headline window: sizes as the oracle's, the device time of the whole marginalisation in the result."""
    win = synth.make_window("C3")
    dense, _, info = solver.marginalize(win)
    marg, keep, idx, m, n = marginalize.pre_marginalize(win, None)
    assert (info["m"], info["n"]) == (m, n) and list(info["keep"]) == keep
    A0, b0 = marginalize.information(win, marg, keep, idx, m, n, None)
    assert np.abs(info["A"] - A0).max() <= 1e-12 * np.abs(A0).max()
    out0 = oracle.schur_prior(A0, b0, m)
    scale = np.abs(out0["Ak"]).max()
    assert np.abs(info["Ak"] - out0["Ak"]).max() <= 1e-9 * scale
    assert np.abs(dense.J.T @ dense.J - out0["Ak"]).max() <= 1e-8 * scale
    print(f"C3 marginalisation: m = {m}, n = {n}, n_full = {info['n_full']}, sweeps {info['eig_sweeps_m']}/{info['eig_sweeps_n']}, "
          f"device {info['ms_device']:.3f} ms, host total {info['ms_total_host']:.3f} ms")

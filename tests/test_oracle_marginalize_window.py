"""Oracle groundwork for row a15 (continued): the marginalisation DRIVER of the reference restated on the flattened window
(oracle/marginalize.py: preMarginalize, the marginalisation blocks of AngularAdjustmentCERESAnalytic::marginalize, the
information assembly) and chained with oracle/marg.hpp; the prior it produces is fed back into the window solve."""
import numpy as np
import pytest

from oracle import marginalize, oracle
from sadvio_b200 import synth


@pytest.fixture(scope="module")
def small():
    win = synth.make_window("small")
    prior, info = marginalize.marginalize_oldest(win)
    return win, prior, info


def test_index_map_follows_the_reference_order(small):
    win, prior, info = small
    idx, m, n = info["idx"], info["m"], info["n"]
    assert idx["f0"] == 0                                           # marginalization.cpp:40-41
    assert m == 15 + 3 * len(info["marg"])                          # pose 6 + v, ba, bg 9 (:45-48) + lonely landmarks (:86-88)
    for k, l in enumerate(info["marg"]):
        assert idx[l] == 15 + 3 * k                                 # :93-98
    assert idx["f1"] == m and n == 15 + 3 * len(info["keep"])       # :101-106
    for k, l in enumerate(info["keep"]):
        assert idx[l] == m + 15 + 3 * k                             # :109-114
    # every kept landmark is seen by frame 0 through both cameras and by another keyframe; lonely ones only by frame 0
    f0 = win.n_frames - 1
    for l in info["keep"]:
        sel = win.obs_lmk == l
        assert np.count_nonzero(sel & (win.obs_frame == f0)) == 2 and np.any(sel & (win.obs_frame != f0))
    for l in info["marg"]:
        assert not np.any((win.obs_lmk == l) & (win.obs_frame != f0))
    # after computeSchurComplement the map is shifted by m (marginalization.cpp:251-256): frame 1 first, then the landmarks
    assert prior.frame == win.n_frames - 2 and prior.frame_col == 0
    assert list(prior.keep_col) == [15 + 3 * k for k in range(len(info["keep"]))]


@pytest.mark.parametrize("kind", [0, 1])
def test_gradient_of_the_marginalised_factor_set(small, kind):
    """b = sum J^T r must be the gradient of 1/2 sum r^2 of the marginalised factors w.r.t. the stacked parameters: checked by
    central differences along random directions (exercises the index map and every block of the assembly); kind 0 = bearing
    factors (AngularAdjustmentCERESAnalytic::marginalize), 1 = pixel factors (BundleAdjustmentCERESAnalytic::marginalize)."""
    win, prior, info = small
    if kind == 1:
        win = synth.make_window("small", factor_kind=1)
        win.has_prior[win.n_frames - 2] = 1                          # a pose prior on frame 1: the pixel optimizer must ignore it
        win.T_prior[win.n_frames - 2] = win.T_f_w[win.n_frames - 2] + 0.01
        win.inf_prior[win.n_frames - 2] = 50.0
        prior, info = marginalize.marginalize_oldest(win)
    idx, m, n, keep, marg = info["idx"], info["m"], info["n"], info["keep"], info["marg"]
    F = win.n_frames
    f0, f1 = F - 1, F - 2
    p = int(np.flatnonzero((win.imu_i == f0) & (win.imu_j == f1))[0])
    pre = oracle.pack_preint(win.imu_dR[p], win.imu_dv[p], win.imu_dp[p], win.imu_cov[p], win.imu_J_dR_bg[p], win.imu_J_dv_ba[p],
                             win.imu_J_dv_bg[p], win.imu_J_dp_ba[p], win.imu_J_dp_bg[p])
    dt = float(win.imu_dt[p])
    wa, wg = 1.0 / (np.sqrt(dt) * float(win.imu_sigma_ba[p])), 1.0 / (np.sqrt(dt) * float(win.imu_sigma_bg[p]))

    def cost(x):
        a, c1 = idx["f0"], idx["f1"]
        params = np.concatenate([x[a:a + 6], x[c1:c1 + 6], x[a + 6:a + 9], x[c1 + 6:c1 + 9], x[a + 9:a + 12], x[a + 12:a + 15]])
        r, _ = oracle.imu_factor_eval(win.T_f_w[f0], win.T_f_w[f1], win.v[f0], win.v[f1], dt, pre, params, jac=False)
        c = 0.5 * r @ r
        rb = np.concatenate([(win.ba[f1] + x[c1 + 9:c1 + 12] - win.ba[f0] - x[a + 9:a + 12]) * wa,
                             (win.bg[f1] + x[c1 + 12:c1 + 15] - win.bg[f0] - x[a + 12:a + 15]) * wg])
        c += 0.5 * rb @ rb
        for l in list(keep) + list(marg):
            for o in np.flatnonzero((win.obs_lmk == l) & (win.obs_frame == f0)):
                cam = int(win.obs_cam[o])
                focal = 0.5 * (win.K[cam][0] + win.K[cam][1])
                if kind == 1:
                    r, _, _ = oracle.reproj_eval(win.obs_uv[o], win.K[cam], win.T_s_f[cam], win.T_f_w[f0], win.lmk_t[l], 1.0,
                                                 dx=x[a:a + 6], dp=x[idx[l]:idx[l] + 3], jac=False)
                else:
                    r, _, _ = oracle.angular_eval(win.obs_bearing[o], win.T_s_f[cam], win.T_f_w[f0], win.lmk_t[l], 1.0 / focal,
                                                  dx=x[a:a + 6], dp=x[idx[l]:idx[l] + 3], jac=False)
                c += 0.5 * r @ r
        for key, f in ((("f0", f0),) if kind == 1 else (("f0", f0), ("f1", f1))):
            if win.has_prior is not None and win.has_prior[f]:
                r, _ = oracle.pose_prior_eval(win.T_f_w[f], win.T_prior[f], win.inf_prior[f], dx=x[idx[key]:idx[key] + 6], jac=False)
                c += 0.5 * r @ r
        return c

    rng = np.random.default_rng(7)
    b = info["b"]
    for _ in range(3):
        d = rng.normal(size=m + n)
        d /= np.linalg.norm(d)
        h = 1e-6
        num = (cost(h * d) - cost(-h * d)) / (2 * h)
        assert abs(num - b @ d) <= 1e-5 * max(1.0, abs(b @ d))


def test_prior_factor_reproduces_the_marginal_information(small):
    win, prior, info = small
    A, b, m = info["A"], info["b"], info["m"]
    assert np.abs(A - A.T).max() <= 1e-9 * np.abs(A).max()
    Amm = 0.5 * (A[:m, :m] + A[:m, :m].T)
    w, V = np.linalg.eigh(Amm)
    assert w[0] > 1e-6                                  # frame 0 is fully constrained (IMU factor, pose prior, 86 bearings)
    Ainv = (V / w) @ V.T
    Ak = A[m:, m:] - A[m:, :m] @ Ainv @ A[m:, :m].T
    bk = b[m:] - A[m:, :m] @ Ainv @ b[:m]
    scale = np.abs(Ak).max()
    assert np.abs(info["Ak"] - Ak).max() <= 1e-9 * scale
    assert np.abs(prior.J.T @ prior.J - Ak).max() <= 1e-8 * scale
    assert np.abs(prior.J.T @ prior.r0 + bk).max() <= 1e-8 * max(1.0, np.abs(bk).max())
    assert prior.J.shape[1] == 15 + 3 * len(info["keep"])


def _next_window_error(win, prior, full, n_fixed):
    from sadvio_b200 import abi  # noqa: F401
    w2 = marginalize.drop_oldest_frame(win, prior)
    w2.n_fixed = n_fixed
    rc, d2, st = oracle.solve_window(w2, nthreads=4)
    assert rc == 0
    f1 = win.n_frames - 2
    new = synth.apply_delta(w2, d2)
    return float(np.abs(new["T_f_w"][f1] - full["T_f_w"][f1]).max())


def test_reference_sign_convention_of_r0(small):
    """REFERENCE QUIRK, reproduced on purpose: b accumulates +J^T r (marginalization.cpp:184) but the prior residual is
    r0 = -Lambda^-1/2 U^T b_k (:527) and the factor evaluates r0 + J dx (marginalization.hpp:148), so the prior's gradient at
    the linearisation point is -b_k, the opposite of the marginal gradient.  It is harmless where the reference uses it
    (marginalisation right after the window solve, b_k ~ 0).  Marginalising at the perturbed initial state instead makes it
    visible: with the sign flipped the shorter window lands 30 x closer to the full-window solution than without any prior,
    with the reference's sign it is pushed away.  This also shows that J, |r0| and the column maps are right."""
    from sadvio_b200 import abi

    win, prior, info = small
    assert np.abs(info["bk"]).max() > 1e3            # far from the optimum: a large marginal gradient
    rc, d_full, _ = oracle.solve_window(win, nthreads=4)
    full = synth.apply_delta(win, d_full)
    flipped = abi.DensePrior(J=prior.J, r0=-prior.r0, frame=prior.frame, frame_col=prior.frame_col, keep_lmk=prior.keep_lmk, keep_col=prior.keep_col)
    e_ref = _next_window_error(win, prior, full, 0)
    e_flip = _next_window_error(win, flipped, full, 0)
    e_none = _next_window_error(win, None, full, 1)   # no prior: the oldest remaining keyframe is held constant instead
    assert e_flip < 0.1 * e_none
    assert e_ref > e_none


def test_marginalising_after_the_solve():
    """The reference's order (slamBiMonoVIO.cpp:570-594): optimise, then marginalise, then solve the next window with the
    prior.  Plumbing check: the prior has full rank up to the 6-dof gauge, the next window accepts it, lowers its cost and
    leaves the oldest remaining keyframe within centimetres of where the previous solve put it.  (Frame 0 is held constant in
    the window solve but is a free block in the marginalisation, AOptimizer.cpp:46-51 vs …Analytic.cpp:504-505, so the prior is
    not expected to be exactly stationary at that point.)"""
    win = synth.make_window("small")
    cfg = oracle.default_config()
    cfg.function_tolerance = 1e-10
    cfg.max_num_iterations = 50
    rc, d, st = oracle.solve_window(win, cfg, nthreads=4)
    assert rc == 0
    new = synth.apply_delta(win, d)
    win.T_f_w, win.lmk_t = new["T_f_w"], new["lmk_t"]
    win.v, win.ba, win.bg = new["v"], new["ba"], new["bg"]
    win.normalise()
    prior, info = marginalize.marginalize_oldest(win)
    assert prior is not None and info["n_full"] >= info["n"] - 6       # at most the 6-dof gauge is dropped
    w2 = marginalize.drop_oldest_frame(win, prior)
    assert w2.n_frames == win.n_frames - 1 and w2.dense_prior is not None
    rc, d2, st2 = oracle.solve_window(w2, cfg, nthreads=4)
    assert rc == 0 and st2["final_cost"] <= st2["initial_cost"]
    assert float(np.abs(d2.dpose[win.n_frames - 2]).max()) < 0.1


def test_sparsified_vio_prior(small):
    """Marginalization::sparsifyVIO restated (oracle/marginalize.py): every factor carries the information of its own
    measurement function under the marginal covariance; the shorter window solves with the sparsified prior."""
    win, prior, info = small
    sp = marginalize.sparsify_vio(win, info)
    n_keep = len(info["keep"])
    assert sp.has_imu_prior and sp.frame == win.n_frames - 2
    assert sp.p2l_lmk.shape == (n_keep,) and sp.p2l_delta.shape == (n_keep, 3) and sp.p2l_sqrt_inf.shape == (n_keep, 9)
    Sigma_k = info["U"] @ np.diag(1.0 / info["Lambda"]) @ info["U"].T        # marginalization.cpp:262
    T = win.T_f_w[win.n_frames - 2].reshape(3, 4)
    for k in (0, n_keep // 2, n_keep - 1):
        S = sp.p2l_sqrt_inf[k].reshape(3, 3)
        assert np.abs(S - S.T).max() < 1e-9 * np.abs(S).max() and np.all(np.linalg.eigvalsh(S) > 0)
        # information of h(x) = R (p + dp) + t + (-R [t]x dw + R dt) under Sigma_k
        c = 15 + 3 * k
        J = np.zeros((3, info["n"]))
        J[:, c:c + 3] = T[:, :3]
        tx = np.array([[0, -T[2, 3], T[1, 3]], [T[2, 3], 0, -T[0, 3]], [-T[1, 3], T[0, 3], 0]])
        J[:, 0:3], J[:, 3:6] = -T[:, :3] @ tx, T[:, :3]
        assert np.abs(S @ S - np.linalg.inv(J @ Sigma_k @ J.T)).max() <= 1e-6 * np.abs(S @ S).max()
        assert np.allclose(sp.p2l_delta[k], T[:, :3] @ win.lmk_t[info["keep"][k]] + T[:, 3])
    S15 = sp.imu_sqrt_inf.reshape(15, 15)
    w15 = np.linalg.eigvalsh(S15)   # positive semi-definite: eigenvalues at or below the reference's threshold are zeroed (:401)
    assert np.abs(S15 - S15.T).max() < 1e-9 * np.abs(S15).max() and w15[0] > -1e-9 * w15[-1] and w15[-1] > 0
    w2 = marginalize.with_sparse_prior(win, sp)
    rc, d2, st = oracle.solve_window(w2, nthreads=4)
    assert rc == 0 and st["final_cost"] <= st["initial_cost"]


def test_sparsified_vo_prior():
    """Marginalization::sparsifyVO restated (oracle/marginalize.py): greedy chain over the coupling of the information blocks,
    unary factor on the landmark of least entropy, relative factors along the chain; the shorter window solves with it."""
    win = synth.make_window("small", vio=False)
    prior, info = marginalize.marginalize_oldest(win)
    assert prior is not None and prior.frame == -1 and info["m"] == 6 + 3 * len(info["marg"]) and info["n"] == 3 * len(info["keep"])
    sp = marginalize.sparsify_vo(win, info)
    keep = list(info["keep"])
    K = len(keep)
    chain = [int(sp.l2l_a[0])] + [int(x) for x in sp.l2l_b]
    # a chain: consecutive links share a landmark, no landmark twice, all of them kept landmarks
    assert np.array_equal(sp.l2l_a[1:], sp.l2l_b[:-1]) and len(set(chain)) == len(chain) <= K and set(chain) <= set(keep)
    assert sp.has_lmk_prior and sp.lmk0 in chain and not sp.has_imu_prior
    Ak = info["Ak"]
    blk = lambda a, b: abs(np.trace(Ak[3 * keep.index(a):3 * keep.index(a) + 3, 3 * keep.index(b):3 * keep.index(b) + 3]))
    mi = np.array([[0.0 if a == b else blk(a, b) for b in keep] for a in keep])
    # the first link is the strongest coupling of all; every later link is the strongest coupling of the chain's tail with a
    # landmark not yet on the chain
    assert blk(chain[0], chain[1]) == mi.max()
    for k in range(1, len(chain) - 1):
        rest = [l for l in keep if l not in chain[:k + 1]]
        assert blk(chain[k], chain[k + 1]) == max(blk(chain[k], l) for l in rest)
    Sigma_k = info["U"] @ np.diag(1.0 / info["Lambda"]) @ info["U"].T
    det = {l: np.linalg.det(Sigma_k[3 * keep.index(l):3 * keep.index(l) + 3, 3 * keep.index(l):3 * keep.index(l) + 3]) for l in chain}
    assert det[sp.lmk0] == min(det.values())                          # entropy is monotone in the determinant
    assert np.allclose(sp.lmk_prior, win.lmk_t[sp.lmk0])
    c0 = 3 * keep.index(sp.lmk0)
    S0 = sp.lmk_sqrt_inf.reshape(3, 3)
    assert np.abs(S0 @ S0 - np.linalg.inv(Sigma_k[c0:c0 + 3, c0:c0 + 3])).max() <= 1e-6 * np.abs(S0 @ S0).max()
    for k in (0, len(chain) // 2, len(chain) - 2):
        a, b = 3 * keep.index(chain[k]), 3 * keep.index(chain[k + 1])
        J = np.zeros((3, info["n"]))
        J[:, a:a + 3], J[:, b:b + 3] = np.eye(3), -np.eye(3)
        S = sp.l2l_sqrt_inf[k].reshape(3, 3)
        assert np.abs(S - S.T).max() < 1e-9 * np.abs(S).max()
        assert np.abs(S @ S - np.linalg.inv(J @ Sigma_k @ J.T)).max() <= 1e-6 * np.abs(S @ S).max()
        assert np.allclose(sp.l2l_delta[k], win.lmk_t[chain[k]] - win.lmk_t[chain[k + 1]])
    w2 = marginalize.with_sparse_prior(win, sp)
    assert w2.sparse_prior.lmk0 >= 0 and w2.sparse_prior.l2l_a.max() < w2.n_lmks
    rc, d2, st = oracle.solve_window(w2, nthreads=4)
    assert rc == 0 and st["final_cost"] <= st["initial_cost"]


def _toy_window():
    """The reference's toy graph (marginalization_test.cpp:26-197) as a flattened window: frame 0 at the origin, frame 1 one metre
    ahead, a stereo pair 0.2 apart on both, landmark 0 seen from frame 0 only, landmarks 1 and 2 from both frames."""
    from sadvio_b200 import abi

    K = np.array([100.0, 100.0, 400.0, 400.0])
    T_left, T_right = np.eye(3, 4), np.eye(3, 4)
    T_right[1, 3] = 0.2                                                     # :47-49
    T_f0, T_f1 = np.eye(3, 4), np.eye(3, 4)
    T_f1[2, 3] = -1.0                                                       # T_w_f1.translation = (0, 0, 1), inverted (:64-66)
    lmk = np.array([[0.5, 0, 2.0], [-1.0, 0, 2.0], [1.0, 0, 2.0]])          # :70, :79, :88
    T_f_w = np.stack([T_f1.reshape(12), T_f0.reshape(12)])                  # newest first
    obs = [(0, 1, 0), (0, 1, 1), (1, 1, 0), (1, 1, 1), (1, 0, 0), (1, 0, 1), (2, 1, 0), (2, 1, 1), (2, 0, 0), (2, 0, 1)]
    uv, bearing = [], []
    for l, f, c in obs:
        Tf, Ts = T_f_w[f].reshape(3, 4), (T_left, T_right)[c]
        pc = Ts[:, :3] @ (Tf[:, :3] @ lmk[l] + Tf[:, 3]) + Ts[:, 3]
        p = np.array([K[0] * pc[0] / pc[2] + K[2], K[1] * pc[1] / pc[2] + K[3]])
        uv.append(p)
        bearing.append(synth.ray_camera(K, p[None])[0])
    o = np.array(obs, dtype=np.int32)
    return abi.Window(vio=False, factor_kind=abi.SDV_FACTOR_ANGULAR, n_fixed=0, T_f_w=T_f_w, T_s_f=np.stack([T_left.reshape(12), T_right.reshape(12)]),
                      K=np.stack([K, K]), lmk_t=lmk, obs_lmk=o[:, 0], obs_frame=o[:, 1], obs_cam=o[:, 2], obs_bearing=np.array(bearing),
                      obs_uv=np.array(uv), has_prior=np.zeros(2, np.uint8), T_prior=np.tile(np.eye(3, 4).reshape(12), (2, 1)),
                      inf_prior=np.zeros((2, 6))).normalise()


def test_reference_toy_graph_through_the_driver():
    """preMargTest / margTest of the reference (marginalization_test.cpp:213-317) through oracle/marginalize.py."""
    win = _toy_window()
    marg, keep, idx, m, n = marginalize.pre_marginalize(win)
    assert (n, m) == (6, 9)                                                 # ASSERT_EQ(_marg._n, 6), (_marg._m, 9)      (:219-220)
    assert marg == [0] and keep == [1, 2]                                   # lmk_to_marg = {lmk_0}, two landmarks kept   (:222-223)
    assert (idx["f0"], idx[0], idx[1], idx[2]) == (0, 6, 9, 12)             # index map before the Schur complement       (:300-303)
    prior, info = marginalize.marginalize_oldest(win)
    assert prior is not None and prior.frame == -1
    assert list(prior.keep_lmk) == [1, 2] and list(prior.keep_col) == [0, 3]  # shifted by m afterwards                   (:308-309)
    Ak = info["Ak"]
    assert Ak.size == 36 and np.linalg.norm(Ak - Ak.T) < 1e-8               #                                              (:312-313)
    assert abs(np.trace(Ak[0:3, 3:6])) > 0                                  # computeOffDiag(lmk_1, lmk_2) > 0            (:315-317)
    # computeSchurComplement refuses fewer than 4 kept parameters (marginalization.cpp:215; margFailTest :321-334 reaches it
    # with an uncorrelated frame, n = 0): here a single kept landmark, n = 3
    sel = win.obs_lmk != 2
    for k in ("obs_lmk", "obs_frame", "obs_cam", "obs_bearing", "obs_uv"):
        setattr(win, k, getattr(win, k)[sel])
    prior, _ = marginalize.marginalize_oldest(win.normalise())
    assert prior is None


def _schur(A, b, m_idx, r_idx):
    Amm = A[np.ix_(m_idx, m_idx)]
    Arm = A[np.ix_(r_idx, m_idx)]
    Ainv = np.linalg.inv(0.5 * (Amm + Amm.T))
    return A[np.ix_(r_idx, r_idx)] - Arm @ Ainv @ Arm.T, b[r_idx] - Arm @ Ainv @ b[m_idx]


def test_chained_marginalisation_conserves_information():
    """The second marginalisation folds the previous prior in as one more block (…Analytic.cpp:631-660, preMarginalize
    :118-142).  With the state untouched in between, marginalising frame A and then frame B must give the information the
    joint elimination of both gives (Schur complements compose) — with the gradient of the first prior entering NEGATED, the
    reference's r0 sign convention (test_reference_sign_convention_of_r0)."""
    win = synth.make_window("small")
    F = win.n_frames
    prior1, info1 = marginalize.marginalize_oldest(win)
    w1 = marginalize.drop_oldest_frame(win, prior1)
    assert w1.dense_prior.frame == w1.n_frames - 1                       # the prior sits on the frame that goes next
    prior2, info2 = marginalize.marginalize_oldest(w1)                   # chained: w1.dense_prior is _marginalization_last
    assert prior2 is not None
    lmk_from = w1.meta["lmk_from"]
    # every landmark of the first prior is a variable of the second marginalisation: marginalised if frame B was its last
    # observer, kept otherwise (also when frame B does not see it at all — the "resurrected" branch)
    for l in w1.dense_prior.keep_lmk:
        assert int(l) in info2["idx"]
    # ---- joint information over (frame A, frame B, frame C, landmarks in `win` numbering)
    w1_plain = marginalize.drop_oldest_frame(win, None)
    marg2, keep2, idx2, m2, n2 = marginalize.pre_marginalize(w1_plain, w1.dense_prior)
    assert (marg2, keep2, idx2, m2, n2) == (info2["marg"], info2["keep"], info2["idx"], info2["m"], info2["n"])
    A2, b2 = marginalize.information(w1_plain, marg2, keep2, idx2, m2, n2, None)      # the factors of step 2 without the prior
    A1, b1 = info1["A"], info1["b"]
    col, nxt = {}, 0

    def var(key, size):
        nonlocal nxt
        if key not in col:
            col[key] = nxt
            nxt += size
        return col[key]

    maps = []
    for idx, names in ((info1["idx"], {"f0": "A", "f1": "B"}), (idx2, {"f0": "B", "f1": "C"})):
        mp = {}
        for k, c in idx.items():
            if k in names:
                mp[c] = (var(names[k], 15), 15)
            else:
                g = int(k) if names["f0"] == "A" else int(lmk_from[int(k)])
                mp[c] = (var(("l", g), 3), 3)
        maps.append(mp)
    N = nxt
    Aj, bj = np.zeros((N, N)), np.zeros(N)
    for (Ax, bx, sign), mp in (((A1, b1, -1.0), maps[0]), ((A2, b2, 1.0), maps[1])):
        loc = np.concatenate([np.arange(c, c + s) for c, (g, s) in sorted(mp.items())])
        glo = np.concatenate([np.arange(g, g + s) for c, (g, s) in sorted(mp.items())])
        Aj[np.ix_(glo, glo)] += Ax[np.ix_(loc, loc)]
        bj[glo] += sign * bx[loc]
    # variables that survive step 2, in the order of the second prior's columns
    keep_cols = np.concatenate([np.arange(col["C"], col["C"] + 15)] + [np.arange(col[("l", int(lmk_from[l]))], col[("l", int(lmk_from[l]))] + 3) for l in keep2])
    gone = np.setdiff1d(np.arange(N), keep_cols)
    Ak_joint, bk_joint = _schur(Aj, bj, gone, keep_cols)
    scale = np.abs(Ak_joint).max()
    assert np.abs(info2["Ak"] - Ak_joint).max() <= 1e-7 * scale
    assert np.abs(info2["bk"] - bk_joint).max() <= 1e-7 * max(1.0, np.abs(bk_joint).max())
    # and the back end's loop goes on: solve, write back, marginalise, drop — until landmarks that entered a prior while other
    # keyframes still saw them are marginalised themselves (their last observer becomes the oldest keyframe)
    from sadvio_b200 import api

    w, from_prior_then_marginalised = marginalize.drop_oldest_frame(w1, prior2), 0
    assert w.n_frames == F - 2
    for _ in range(4):
        rc, d, st = oracle.solve_window(w, nthreads=4)
        assert rc == 0 and st["final_cost"] <= st["initial_cost"]
        api.write_back(w, d, True)
        last_keep = set(int(l) for l in w.dense_prior.keep_lmk)
        prior, info = marginalize.marginalize_oldest(w)
        assert prior is not None and np.abs(info["Ak"] - info["Ak"].T).max() <= 1e-9 * np.abs(info["Ak"]).max()
        from_prior_then_marginalised += len(last_keep & set(info["marg"]))
        w = marginalize.drop_oldest_frame(w, prior)
    assert w.n_frames == F - 6 and from_prior_then_marginalised > 0


def test_pixel_optimizer_marginalisation_feeds_the_next_window():
    """BundleAdjustmentCERESAnalytic::marginalize (pixel factors, sigma = 1): prior reproduces the marginal information, the next
    window solves with it."""
    win = synth.make_window("small", factor_kind=1)
    rc, d, st = oracle.solve_window(win, nthreads=4)
    from sadvio_b200 import api
    api.write_back(win, d, True)
    prior, info = marginalize.marginalize_oldest(win)
    assert prior is not None and prior.J.shape[1] == 15 + 3 * len(info["keep"])
    scale = np.abs(info["Ak"]).max()
    assert np.abs(prior.J.T @ prior.J - info["Ak"]).max() <= 1e-8 * scale
    w2 = marginalize.drop_oldest_frame(win, prior)
    assert w2.factor_kind == 1
    rc, d2, st2 = oracle.solve_window(w2, nthreads=4)
    assert rc == 0 and st2["final_cost"] <= st2["initial_cost"]

"""TEST INFRASTRUCTURE (oracle/README.md) — the checker of sdv_marginalize.

Restatement, on the flattened window of the C ABI, of the marginalisation the back end runs before every window solve
(first and chained: a dense prior already attached to the window is `_marginalization_last`), VIO and VO, for both optimizers:

    Marginalization::preMarginalize              cpp/src/optimizers/marginalization.cpp:23-143
    AngularAdjustmentCERESAnalytic::marginalize  cpp/src/optimizers/AngularAdjustmentCERESAnalytic.cpp:488-739   (factor_kind 0)
    BundleAdjustmentCERESAnalytic::marginalize   cpp/src/optimizers/BundleAdjustmentCERESAnalytic.cpp:431-660    (factor_kind 1)
    computeInformationAndGradient                cpp/src/optimizers/marginalization.cpp:145-211
    computeSchurComplement / rankRevealling / computeJacobiansAndResiduals   (oracle/marg.hpp)
    sparsifyVIO / sparsifyVO                     cpp/src/optimizers/marginalization.cpp:362-514

frame0 = the oldest keyframe of the window (last index: frames are ordered newest -> oldest, amap.h:28-32), frame1 = the next
one.  The factors come from the oracle's functor restatements (oracle.angular_eval, reproj_eval, imu_factor_eval,
pose_prior_eval)."""
from __future__ import annotations

import numpy as np

from oracle import oracle
from sadvio_b200 import abi


def pre_marginalize(win: abi.Window, last: abi.DensePrior | None = None):
    """marginalization.cpp:23-143 for point landmarks: returns (marg, keep, idx) — landmark ids to marginalise / to keep (in
    landmark order, as getLandmarks() is walked) and the parameter index map {('f0'|'f1'|landmark id): first column}.
    `last` is the previous marginalisation (`_marginalization_last`) in this window's indices: its kept landmarks carry the
    hasPrior() flag (:74) and are appended to the kept set when frame 0 does not bring them in itself (:118-142; the
    outlier branch :124-127 cannot occur here, outliers never reach a flattened window)."""
    F = win.n_frames
    f0 = F - 1
    marg, keep = [], []
    with_prior = set(int(l) for l in last.keep_lmk) if last is not None else set()
    lm_of_f0 = np.unique(win.obs_lmk[win.obs_frame == f0])
    for l in lm_of_f0:
        sel = win.obs_lmk == l
        num_cam = int(np.count_nonzero(sel & (win.obs_frame == f0)))      # :58-71
        lonely = not np.any(sel & (win.obs_frame != f0))
        if num_cam != 2 and int(l) not in with_prior:                      # no stereo pair, no prior: ignored (:74-77)
            continue
        (marg if lonely else keep).append(int(l))                          # :80-89
    idx = {"f0": 0}
    last_idx = 6 + (9 if win.vio else 0)                                       # :40-48
    for l in marg:                                                         # :93-98
        idx[l] = last_idx
        last_idx += 3
    m = last_idx
    n = 0
    if win.vio:                                                            # :101-106
        idx["f1"] = last_idx
        last_idx += 15
        n += 15
    for l in keep:                                                         # :109-114
        idx[l] = last_idx
        last_idx += 3
        n += 3
    for l in ([int(q) for q in last.keep_lmk] if last is not None else []):   # "resurrected" landmarks, :118-139
        if l not in idx:
            keep.append(l)
            idx[l] = last_idx
            last_idx += 3
            n += 3
    return marg, keep, idx, m, n


def information(win: abi.Window, marg, keep, idx, m, n, last: abi.DensePrior | None = None):
    """The marginalisation blocks of AngularAdjustmentCERESAnalytic.cpp:504-687 (bearing factors) or
    BundleAdjustmentCERESAnalytic.cpp:446-617 (pixel factors; `win.factor_kind`) evaluated at the current state and accumulated as in
    marginalization.cpp:145-211.  Returns A [(m+n)^2], b [m+n].  `last`: the previous prior, added as one more block
    (…Analytic.cpp:631-660) — MarginalizationFactor at dx = 0, i.e. residual r0 and the column slices of its J."""
    F = win.n_frames
    f0, f1 = F - 1, F - 2
    N = m + n
    A, b = np.zeros((N, N)), np.zeros(N)

    def add(blocks, r):
        for i, (ci, Ji) in enumerate(blocks):
            for cj, Jj in blocks[i:]:
                A[ci:ci + Ji.shape[1], cj:cj + Jj.shape[1]] += Ji.T @ Jj
                if cj != ci:
                    A[cj:cj + Jj.shape[1], ci:ci + Ji.shape[1]] = A[ci:ci + Ji.shape[1], cj:cj + Jj.shape[1]].T
            b[ci:ci + Ji.shape[1]] += Ji.T @ r

    if win.vio:
        p = int(np.flatnonzero((win.imu_i == f0) & (win.imu_j == f1))[0])
        pre = oracle.pack_preint(win.imu_dR[p], win.imu_dv[p], win.imu_dp[p], win.imu_cov[p], win.imu_J_dR_bg[p], win.imu_J_dv_ba[p],
                                 win.imu_J_dv_bg[p], win.imu_J_dp_ba[p], win.imu_J_dp_bg[p])
        r, J = oracle.imu_factor_eval(win.T_f_w[f0], win.T_f_w[f1], win.v[f0], win.v[f1], float(win.imu_dt[p]), pre)
        # parameter blocks in the order of …Analytic.cpp:522-537: pose0, pose1, v0, v1, dba0, dbg0
        add([(idx["f0"], J[:, 0:6]), (idx["f1"], J[:, 6:12]), (idx["f0"] + 6, J[:, 12:15]), (idx["f1"] + 6, J[:, 15:18]),
             (idx["f0"] + 9, J[:, 18:21]), (idx["f0"] + 12, J[:, 21:24])], r)
        # IMUBiasFactor (residuals.hpp:252-296), blocks ba0, bg0, ba1, bg1 (…Analytic.cpp:549-556)
        dt = float(win.imu_dt[p])
        wa, wg = 1.0 / (np.sqrt(dt) * float(win.imu_sigma_ba[p])), 1.0 / (np.sqrt(dt) * float(win.imu_sigma_bg[p]))
        rb = np.concatenate([(win.ba[f1] - win.ba[f0]) * wa, (win.bg[f1] - win.bg[f0]) * wg])
        Za, Zg = np.zeros((6, 3)), np.zeros((6, 3))
        Za[0:3] = np.eye(3) * wa
        Zg[3:6] = np.eye(3) * wg
        add([(idx["f0"] + 9, -Za), (idx["f0"] + 12, -Zg), (idx["f1"] + 9, Za), (idx["f1"] + 12, Zg)], rb)
    pixel = win.factor_kind == abi.SDV_FACTOR_PIXEL
    for l in list(keep) + list(marg):
        for o in np.flatnonzero((win.obs_lmk == l) & (win.obs_frame == f0)):
            c = int(win.obs_cam[o])
            if pixel:                                                      # BundleAdjustmentCERESAnalytic.cpp:512-571, sigma = 1
                r, J6, J3 = oracle.reproj_eval(win.obs_uv[o], win.K[c], win.T_s_f[c], win.T_f_w[f0], win.lmk_t[l])
            else:                                                          # AngularAdjustment…Analytic.cpp:565-629, sigma = 1 / focal
                focal = 0.5 * (win.K[c][0] + win.K[c][1])
                r, J6, J3 = oracle.angular_eval(win.obs_bearing[o], win.T_s_f[c], win.T_f_w[f0], win.lmk_t[l], 1.0 / focal)
            add([(idx["f0"], J6), (idx[l], J3)], r)
    if last is not None and len(last.keep_lmk):                            # …Analytic.cpp:631-660
        blocks = []
        if last.frame >= 0:                                                # _marginalization_last->_frame_to_keep is frame 0 now
            assert last.frame == f0, "the previous prior must sit on the frame that is marginalised now"
            c = last.frame_col
            blocks += [(idx["f0"], last.J[:, c:c + 6]), (idx["f0"] + 6, last.J[:, c + 6:c + 9]), (idx["f0"] + 9, last.J[:, c + 9:c + 12]),
                       (idx["f0"] + 12, last.J[:, c + 12:c + 15])]
        for l, c in zip(last.keep_lmk, last.keep_col):
            if c >= 0:                                                     # marginalization.hpp:139
                blocks.append((idx[int(l)], last.J[:, c:c + 3]))
        add(blocks, last.r0)
    # pose priors: the angular optimizer adds frame 0's and frame 1's (…Analytic.cpp:664-687), the pixel one only frame 0's
    # (BundleAdjustmentCERESAnalytic.cpp:606-617)
    for key, f in ((("f0", f0),) if pixel else (("f0", f0), ("f1", f1))):
        if win.has_prior is not None and win.has_prior[f] and key in idx:
            r, J = oracle.pose_prior_eval(win.T_f_w[f], win.T_prior[f], win.inf_prior[f])
            add([(idx[key], J)], r)
    return A, b


def marginalize_oldest(win: abi.Window, eps: float = 1e-12):
    """Returns (prior, info): prior = abi.DensePrior over (frame1's 15 parameters, kept landmarks) expressed for the window
    WITHOUT its oldest keyframe, or None when the reference's marginalize() returns false; info = the intermediate results.
    A dense prior already attached to the window is the previous marginalisation and is folded in (chained marginalisation)."""
    last = win.dense_prior
    marg, keep, idx, m, n = pre_marginalize(win, last)
    A, b = information(win, marg, keep, idx, m, n, last)
    out = oracle.schur_prior(A, b, m, eps)
    if out is None:                                                        # …Analytic.cpp:690-695
        return None, {"marg": marg, "keep": keep, "idx": idx, "m": m, "n": n, "A": A, "b": b}
    frame = win.n_frames - 2 if win.vio else -1
    first = 15 if win.vio else 0
    prior = abi.DensePrior(J=np.ascontiguousarray(out["J"]), r0=out["r0"].copy(), frame=frame, frame_col=0,
                           keep_lmk=np.asarray(keep, dtype=np.int32),
                           keep_col=np.asarray([first + 3 * k for k in range(len(keep))], dtype=np.int32))
    out.update({"marg": marg, "keep": keep, "idx": idx, "m": m, "n": n, "A": A, "b": b})
    return prior, out


def drop_oldest_frame(win: abi.Window, prior: abi.DensePrior | None) -> abi.Window:
    """The window the back end optimises next: the oldest keyframe is gone (with its observations, its IMU pair and the
    landmarks only it saw), the marginal prior takes its place; no keyframe is held constant any more."""
    import copy

    F = win.n_frames
    f0 = F - 1
    keep_obs = win.obs_frame != f0
    lm_alive = np.zeros(win.n_lmks, dtype=bool)
    lm_alive[np.unique(win.obs_lmk[keep_obs])] = True
    if prior is not None:
        lm_alive[prior.keep_lmk] = True       # a kept landmark nobody else observes stays as a prior-only parameter block
    remap = np.cumsum(lm_alive) - 1
    w = copy.copy(win)
    w.meta = {"lmk_from": np.flatnonzero(lm_alive)}     # landmark l of the shorter window is landmark lmk_from[l] of `win`
    w.n_fixed = 0
    w.T_f_w = win.T_f_w[:f0].copy()
    for k in ("v", "ba", "bg", "has_imu", "has_prior", "T_prior", "inf_prior"):
        a = getattr(win, k)
        setattr(w, k, None if a is None else a[:f0].copy())
    w.lmk_t = win.lmk_t[lm_alive].copy()
    w.obs_lmk = remap[win.obs_lmk[keep_obs]].astype(np.int32)
    w.obs_frame = win.obs_frame[keep_obs].copy()
    w.obs_cam = win.obs_cam[keep_obs].copy()
    for k in ("obs_bearing", "obs_uv", "obs_sigma"):
        a = getattr(win, k)
        setattr(w, k, None if a is None else a[keep_obs].copy())
    if win.imu_i is not None:
        keep_imu = (win.imu_i != f0) & (win.imu_j != f0)
        for k in ("imu_i", "imu_j", "imu_dt", "imu_dR", "imu_dv", "imu_dp", "imu_cov", "imu_J_dR_bg", "imu_J_dv_ba", "imu_J_dv_bg",
                  "imu_J_dp_ba", "imu_J_dp_bg", "imu_sigma_ba", "imu_sigma_bg"):
            setattr(w, k, getattr(win, k)[keep_imu].copy())
    w.dense_prior = None
    if prior is not None:
        assert np.all(lm_alive[prior.keep_lmk])
        w.dense_prior = abi.DensePrior(J=prior.J, r0=prior.r0, frame=prior.frame, frame_col=prior.frame_col,
                                       keep_lmk=remap[prior.keep_lmk].astype(np.int32), keep_col=prior.keep_col)
    return w.normalise()


def _sym_sqrt(M, eps):
    """V diag(sqrt(max(w, 0 if w <= eps))) V^T — the SelfAdjointEigenSolver idiom of marginalization.cpp:379-385."""
    w, V = np.linalg.eigh(0.5 * (M + M.T))
    return (V * np.sqrt(np.where(w > eps, w, 0.0))) @ V.T


def sparsify_vio(win: abi.Window, info: dict, eps: float = 1e-12) -> abi.SparsePrior:
    """Marginalization::sparsifyVIO (marginalization.cpp:362-411): the dense marginal over (frame1, kept landmarks) is replaced by
    one absolute factor on frame1 (IMUPriordx) and one relative frame-to-landmark factor per kept landmark
    (PoseToLandmarkFactor), each with the information of its own measurement function under the marginal covariance
    U Sigma U^T.  Landmark indices refer to `win`; the prior values of the frame are its current state, as the window wiring
    does (AngularAdjustmentCERESAnalytic.cpp:391-397)."""
    f1 = win.n_frames - 2
    n, U, Sigma = info["n"], info["U"], 1.0 / info["Lambda"]
    T = win.T_f_w[f1].reshape(3, 4)
    R, t = T[:, :3], T[:, 3]
    t_skew = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    keep = info["keep"]
    deltas, sqrts = [], []
    for k, l in enumerate(keep):
        c = 15 + 3 * k                                   # the index map after the shift by m (marginalization.cpp:251-256)
        J = np.zeros((3, n))
        J[:, c:c + 3] = R                                # :372
        J[:, 0:3] = -R @ t_skew                          # :373
        J[:, 3:6] = R                                    # :374
        Jt = J @ U
        inf = np.linalg.inv(Jt @ np.diag(Sigma) @ Jt.T)  # :377
        sqrts.append(_sym_sqrt(inf, eps).reshape(9))
        deltas.append(R @ win.lmk_t[l] + t)              # t_f_lmk, :387
    J = np.zeros((15, n))
    J[:, 0:15] = np.eye(15)                              # :393
    J[0:3, 0:3] = R                                      # :394
    J[0:3, 3:6] = R                                      # :395  (as written in the reference)
    J[3:6, 3:6] = R                                      # :396
    Jt = J @ U
    inf15 = np.linalg.inv(Jt @ np.diag(Sigma) @ Jt.T)
    return abi.SparsePrior(has_imu_prior=True, frame=f1, T_prior=win.T_f_w[f1].copy(), v_prior=win.v[f1].copy(), ba_prior=win.ba[f1].copy(),
                           bg_prior=win.bg[f1].copy(), imu_sqrt_inf=_sym_sqrt(inf15, eps).reshape(225),
                           p2l_lmk=np.asarray(keep, dtype=np.int32), p2l_delta=np.asarray(deltas), p2l_sqrt_inf=np.asarray(sqrts))


def sparsify_vo(win: abi.Window, info: dict, eps: float = 1e-12) -> abi.SparsePrior:
    """Marginalization::sparsifyVO (marginalization.cpp:410-514): the dense marginal over the kept landmarks is replaced by a
    chain — the kept landmarks are re-ordered greedily by the coupling |trace(Ak_kl)| of their information blocks
    (computeOffDiag, :304-316), the landmark with the smallest entropy (:267-274) gets a unary Landmark3DPrior and every
    consecutive pair of the chain a LandmarkToLandmarkFactor, each with the information of its own measurement function under
    the marginal covariance.  Landmarks the greedy walk never reaches drop out of the prior, as in the reference (:462-463).
    Landmark indices refer to `win`."""
    n, U, Sigma, Ak = info["n"], info["U"], 1.0 / info["Lambda"], info["Ak"]
    keep = list(info["keep"])
    K = len(keep)
    if n == 0 or K < 2:
        return None
    first = 15 if win.vio else 0
    col = {l: first + 3 * k for k, l in enumerate(keep)}                  # the index map after the shift by m (:251-256)
    mi = np.zeros((K, K))
    for k in range(K):                                                    # :420-431
        for l in range(K):
            if k == l or mi[k, l] != 0:
                continue
            mi[k, l] = mi[l, k] = abs(np.trace(Ak[col[keep[k]]:col[keep[k]] + 3, col[keep[l]]:col[keep[l]] + 3]))
    # Eigen's maxCoeff visits a column-major matrix column by column and keeps the first maximum (:439)
    flat = int(np.argmax(mi.T.reshape(-1)))
    max_col, max_row = divmod(flat, K)
    order = [keep[max_row], keep[max_col]]                                # :440-441
    mi[:, max_row] = 0                                                    # :444-446
    mi[max_row, :] = 0
    mi[:, max_col] = 0
    cur = max_col
    while True:                                                           # :451-459
        c = int(np.argmax(mi[cur]))
        if mi[cur, c] == 0:
            break
        order.append(keep[c])
        mi[cur, :] = 0
        mi[:, c] = 0
        cur = c
    Sigma_k = U @ np.diag(Sigma) @ U.T                                    # :262
    # :273 — std::pow(2 pi e, size / 2) with the INTEGER division the reference writes (3 / 2 = 1)
    ent = [np.log((2 * np.pi * np.e) ** (3 // 2) * np.linalg.det(Sigma_k[col[l]:col[l] + 3, col[l]:col[l] + 3])) for l in order]
    with_prior = order[int(np.argmin(ent))]                               # :469-470

    def sqrt_inf_of(J):                                                   # :481-487, :501-507
        Jt = J @ U
        w, V = np.linalg.eigh(Jt @ np.diag(Sigma) @ Jt.T)
        s = np.where(w > eps, 1.0 / np.where(w > eps, w, 1.0), 0.0)
        return (V * np.sqrt(s)) @ V.T

    J = np.zeros((3, n))
    J[:, col[with_prior]:col[with_prior] + 3] = np.eye(3)                 # :477-478
    S0 = sqrt_inf_of(J)
    a, b, deltas, sqrts = [], [], [], []
    for k in range(len(order) - 1):                                       # :493-511
        lk, lk1 = order[k], order[k + 1]
        J = np.zeros((3, n))
        J[:, col[lk]:col[lk] + 3] = np.eye(3)
        J[:, col[lk1]:col[lk1] + 3] = -np.eye(3)
        sqrts.append(sqrt_inf_of(J).reshape(9))
        deltas.append(win.lmk_t[lk] - win.lmk_t[lk1])
        a.append(lk)
        b.append(lk1)
    return abi.SparsePrior(has_lmk_prior=True, lmk0=int(with_prior), lmk_prior=win.lmk_t[with_prior].copy(), lmk_sqrt_inf=S0.reshape(9),
                           l2l_a=np.asarray(a, dtype=np.int32), l2l_b=np.asarray(b, dtype=np.int32), l2l_delta=np.asarray(deltas),
                           l2l_sqrt_inf=np.asarray(sqrts))


def with_sparse_prior(win: abi.Window, sp: abi.SparsePrior) -> abi.Window:
    """drop_oldest_frame + the sparsified prior (landmark indices remapped to the shorter window)."""
    import copy

    w2 = drop_oldest_frame(win, None)
    keep_obs = win.obs_frame != win.n_frames - 1
    alive = np.zeros(win.n_lmks, dtype=bool)
    alive[np.unique(win.obs_lmk[keep_obs])] = True
    remap = np.cumsum(alive) - 1
    sp2 = copy.copy(sp)
    if sp.p2l_lmk is not None and len(sp.p2l_lmk):
        sp2.p2l_lmk = remap[sp.p2l_lmk].astype(np.int32)
    if sp.has_lmk_prior:
        sp2.lmk0 = int(remap[sp.lmk0])
    if sp.l2l_a is not None and len(sp.l2l_a):
        sp2.l2l_a, sp2.l2l_b = remap[sp.l2l_a].astype(np.int32), remap[sp.l2l_b].astype(np.int32)
    w2.sparse_prior = sp2
    return w2.normalise()

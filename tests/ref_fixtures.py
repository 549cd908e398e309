"""Windows re-stating the reference's own end-to-end optimizer tests (cpp/tests/imu_test.cpp), flattened to abi.Window.

Used by the oracle tests (CPU) and by the GPU parity tests, so both sides are checked against the reference's
published tolerances on the reference's own fixtures.
"""
import numpy as np

from oracle import oracle as orc
from sadvio_b200 import abi

GYR_NOISE = (0.5 * np.pi) / (180 * 60)   # imu_test.cpp:64-68
BGYR_NOISE = 1.9393e-05
ACC_NOISE = 0.1 / 60
BACC_NOISE = 3.0e-3

R_I_F = np.array([[0.38001193, 0.16469125, 0.91020202], [0.03067918, -0.9857245, 0.16554758], [0.92447267, -0.0349858, -0.37963966]])


def eta(rate):
    return (np.array([GYR_NOISE] * 3 + [ACC_NOISE] * 3) ** 2) * rate


def _imu_window(kf, cur, T_kf, T_cur, v_kf, v_cur, prior_kf, prior_cur, dt):
    """2-keyframe IMU-only VIO window; frames newest -> oldest = [cur, kf]."""
    g = lambda s, n: orc.imu_get(s, n)
    return abi.Window(
        vio=True, factor_kind=abi.SDV_FACTOR_ANGULAR, n_fixed=0,
        T_f_w=np.stack([T_cur, T_kf]), T_s_f=np.eye(3, 4).reshape(1, 12), K=np.array([[100.0, 100.0, 400.0, 400.0]]),
        lmk_t=np.zeros((0, 3)), obs_lmk=np.zeros(0, np.int32), obs_frame=np.zeros(0, np.int32), obs_cam=np.zeros(0, np.int32),
        obs_bearing=np.zeros((0, 3)), obs_uv=np.zeros((0, 2)),
        v=np.stack([v_cur, v_kf]), ba=np.stack([g(cur, "ba"), g(kf, "ba")]), bg=np.stack([g(cur, "bg"), g(kf, "bg")]),
        has_imu=np.ones(2, np.uint8), has_prior=np.ones(2, np.uint8), T_prior=np.stack([prior_cur, prior_kf]),
        inf_prior=np.full((2, 6), 100.0),
        imu_i=np.array([1], np.int32), imu_j=np.array([0], np.int32), imu_dt=np.array([dt]),
        imu_dR=g(cur, "dR").reshape(1, 9), imu_dv=g(cur, "dv").reshape(1, 3), imu_dp=g(cur, "dp").reshape(1, 3),
        imu_cov=g(cur, "Sigma").reshape(1, 81), imu_J_dR_bg=g(cur, "J_dR_bg").reshape(1, 9),
        imu_J_dv_ba=g(cur, "J_dv_ba").reshape(1, 9), imu_J_dv_bg=g(cur, "J_dv_bg").reshape(1, 9),
        imu_J_dp_ba=g(cur, "J_dp_ba").reshape(1, 9), imu_J_dp_bg=g(cur, "J_dp_bg").reshape(1, 9),
        imu_sigma_ba=np.array([BACC_NOISE]), imu_sigma_bg=np.array([BGYR_NOISE]),
    ).normalise()


def free_fall_window():
    """imu_test.cpp:363-487 — 1000 steps of rotated free fall, then a perturbed 2-KF inertial optimisation."""
    U, _, Vt = np.linalg.svd(R_I_F)  # Affine3d::rotation() == polar factor of the 8-digit matrix
    T_i_f = np.eye(4)
    T_i_f[:3, :3] = U @ Vt
    T_i_f[:3, 3] = 1.0
    T_f_i = np.linalg.inv(T_i_f)
    acc = T_i_f[:3, :3].T @ np.array([0, 0, 10.81])
    gyr = np.zeros(3)
    kf = orc.imu_state(acc, gyr, T_f_w=T_f_i[:3].reshape(12), is_kf=True)
    cur = kf
    for _ in range(1000):
        cur = orc.process_imu(cur, np.zeros(3), np.zeros(3), 0.001, eta(1000.0), 1000.0, acc, gyr)
    T_cur = np.vstack([orc.imu_get(cur, "T_f_w").reshape(3, 4), [0, 0, 0, 1]])
    # perturbation, imu_test.cpp:466-470
    err = np.eye(4)
    err[:3, 3] = [0.1, 0.05, -0.01]
    T_cur_pert = T_cur @ err
    v_cur = orc.imu_get(cur, "v") + np.array([0.04, 0.02, -0.02])
    win = _imu_window(kf, cur, T_f_i[:3].reshape(12), T_cur_pert[:3].reshape(12), orc.imu_get(kf, "v"), v_cur,
                      T_f_i[:3].reshape(12), T_cur[:3].reshape(12), 1.0)
    win.meta = dict(T_i_f=T_i_f)
    return win


def bias_window():
    """imu_test.cpp:545-568 — biases (0.5,1,1)/(0.1,0.3,0.1) cancel the measurement exactly: they must not move."""
    acc, gyr = np.array([0.5, 1.0, 10.81]), np.array([0.1, 0.3, 0.1])
    ba, bg = np.array([0.5, 1.0, 1.0]), np.array([0.1, 0.3, 0.1])
    imu0 = orc.imu_state(acc, gyr, ba=ba, bg=bg, is_kf=True)
    imu1 = orc.process_imu(imu0, ba, bg, 0.5, eta(200.0), 200.0, acc, gyr)
    I = np.eye(3, 4).reshape(12)
    win = _imu_window(imu0, imu1, I, I, orc.imu_get(imu0, "v"), orc.imu_get(imu1, "v"), I, I, 0.5)
    win.meta = dict(ba=ba, bg=bg)
    return win


def check_free_fall(win, d):
    """Assertions of imu_test.cpp:485-487 on the updated state."""
    from sadvio_b200.synth import apply_delta
    new = apply_delta(win, d)
    T = np.vstack([new["T_f_w"][0].reshape(3, 4), [0, 0, 0, 1]])
    T_w_f = np.linalg.inv(T)
    assert np.linalg.norm(T_w_f[:3, 3] - [1, 1, 1.5]) < 1e-2
    assert np.linalg.norm(new["v"][0] - [0, 0, 1]) < 1e-2
    assert abs((T_w_f[:3, :3].T @ win.meta["T_i_f"][:3, :3]).trace() - 3) < 1e-5


def check_bias(win, d):
    """Assertions of imu_test.cpp:566-567."""
    from sadvio_b200.synth import apply_delta
    new = apply_delta(win, d)
    assert np.linalg.norm(new["bg"][1] - win.meta["bg"]) < 1e-5
    assert np.linalg.norm(new["ba"][1] - win.meta["ba"]) < 1e-5


# ----------------------------------------------------------------------------------------------------------------
# imu_test.cpp:885-945 — the bias-estimation run of simuEuroc: 30 keyframes 0.5 s apart on the EuRoC ground truth,
# IMU samples differentiated from it and corrupted by constant biases, a pose prior on every keyframe,
# localMapVIOptimization(local_map, 1) after every new keyframe; the newest keyframe's biases must be the true ones.
# ----------------------------------------------------------------------------------------------------------------
EUROC_BA = np.array([0.1, 0.2, -0.1])      # :892
EUROC_BG = np.array([0.4, -0.2, 0.01])     # :893
G_W = np.array([0.0, 0.0, -9.81])


class OracleOptimizer:
    """localMapVIOptimization through the CPU oracle with the host mirror's own write-back (sadvio_b200.api.write_back): the
    checker's counterpart of api.B200Optimizer for tests that drive a sequence of solves."""

    def __init__(self, nthreads=1):
        self.nthreads, self.last_stats = nthreads, None

    def localMapVIOptimization(self, win, fixed_frame_number=0):  # noqa: N802
        from sadvio_b200 import api
        win.n_fixed, win.vio = int(fixed_frame_number), True
        rc, d, st = orc.solve_window(win, nthreads=self.nthreads)
        self.last_stats = st
        api.write_back(win, d, True)
        return True


def _euroc_samples():
    """meas_vec / pose_vec / vel_vec / ts_vec of imu_test.cpp:741-757 from the committed slice of euroc_gt.csv
    (tests/golden/make_euroc_slice.py, read by synth.load_euroc_slice)."""
    from sadvio_b200 import synth
    ts, R, p, v = synth.load_euroc_slice()                                    # read_line_euroc, :9-39
    n = ts.size
    acc, gyr = np.zeros((n - 1, 3)), np.zeros((n - 1, 3))
    for k in range(n - 1):
        dt = ts[k + 1] - ts[k]
        acc[k] = (1 / dt) * (R[k + 1].T @ (v[k + 1] - v[k])) - R[k + 1].T @ G_W   # :745
        gyr[k] = (1 / dt) * orc.log_so3(R[k].T @ R[k + 1])                        # :746
    ts_ns = (ts * 1e9).astype(np.uint64)                                      # ts_vec (:752) through Frame::init's integer ns
    return acc, gyr, R[:-1], p[:-1], v[:-1], ts[:-1] * 1e9, ts_ns[:-1]


def _T_f_w(R_w_f, t_w_f):
    return np.hstack([R_w_f.T, (-R_w_f.T @ t_w_f)[:, None]]).reshape(12)


def _kf_window(kfs, sigma_ba, sigma_bg):
    """Flatten the keyframe records (oldest -> newest) the way getLastNFramesIn does: newest first."""
    fr = kfs[::-1]
    F = len(fr)
    g = orc.imu_get
    pairs = [(j + 1, j) for j in range(F - 1) if (fr[j]["ts"] - fr[j + 1]["ts"]) * 1e-9 <= 1]      # AOptimizer.cpp:62-73
    col = lambda name, k: np.stack([g(fr[j]["imu"], name) for _, j in pairs]).reshape(len(pairs), k)
    return abi.Window(
        vio=True, factor_kind=abi.SDV_FACTOR_ANGULAR, n_fixed=1,
        T_f_w=np.stack([f["T_f_w"] for f in fr]), T_s_f=np.eye(3, 4).reshape(1, 12), K=np.array([[100.0, 100.0, 400.0, 400.0]]),
        lmk_t=np.zeros((0, 3)), obs_lmk=np.zeros(0, np.int32), obs_frame=np.zeros(0, np.int32), obs_cam=np.zeros(0, np.int32),
        obs_bearing=np.zeros((0, 3)), obs_uv=np.zeros((0, 2)),
        v=np.stack([g(f["imu"], "v") for f in fr]), ba=np.stack([g(f["imu"], "ba") for f in fr]),
        bg=np.stack([g(f["imu"], "bg") for f in fr]),
        has_imu=np.ones(F, np.uint8), has_prior=np.ones(F, np.uint8), T_prior=np.stack([f["T_prior"] for f in fr]),
        inf_prior=np.full((F, 6), 100.0),
        imu_i=np.array([i for i, _ in pairs], np.int32), imu_j=np.array([j for _, j in pairs], np.int32),
        imu_dt=np.array([(fr[j]["ts"] - fr[i]["ts"]) * 1e-9 for i, j in pairs]),
        imu_dR=col("dR", 9), imu_dv=col("dv", 3), imu_dp=col("dp", 3), imu_cov=col("Sigma", 81), imu_J_dR_bg=col("J_dR_bg", 9),
        imu_J_dv_ba=col("J_dv_ba", 9), imu_J_dv_bg=col("J_dv_bg", 9), imu_J_dp_ba=col("J_dp_ba", 9), imu_J_dp_bg=col("J_dp_bg", 9),
        imu_sigma_ba=np.full(len(pairs), sigma_ba), imu_sigma_bg=np.full(len(pairs), sigma_bg),
    ).normalise(), pairs


def euroc_bias_run(optimizer, n_kf=30, dt_kf=0.5):
    """Drives `optimizer.localMapVIOptimization(window, 1)` exactly as imu_test.cpp:885-942 does; returns the keyframe records
    (oldest first; record["imu"] is the keyframe's IMU object as an oracle state vector) and the per-solve statistics."""
    acc, gyr, R, p, v, ts_f, ts_ns = _euroc_samples()
    bgyr_noise, acc_noise, bacc_noise, rate = 1.9393e-03, 3.0e-2, 3.0e-2, 200.0        # :884-886
    eta6 = (np.array([GYR_NOISE] * 3 + [acc_noise] * 3) ** 2) * rate
    T0 = _T_f_w(R[0], p[0])
    imu0 = orc.imu_state(acc[0] + EUROC_BA, gyr[0] + EUROC_BG, T_f_w=T0, v=v[0], is_kf=True)  # :896-906
    kfs = [dict(T_f_w=T0.copy(), T_prior=T0.copy(), imu=imu0, ts=int(ts_ns[0]))]
    tsp, last, last_ts = ts_f[0], imu0, int(ts_ns[0])
    stats = []
    for i in range(1, acc.shape[0]):
        dt_vote = (ts_f[i] - tsp) * 1e-9                                                     # :911
        kf_imu = kfs[-1]["imu"]
        cur = orc.process_imu(last, orc.imu_get(kf_imu, "ba"), orc.imu_get(kf_imu, "bg"), (int(ts_ns[i]) - last_ts) * 1e-9, eta6, rate,
                              acc[i] + EUROC_BA, gyr[i] + EUROC_BG)                          # :914-921
        T = _T_f_w(R[i], p[i])
        cur[15:27] = T                                                                      # :927 (ground-truth pose)
        cur[12:15] = v[i]                                                                   # :928
        last, last_ts = cur, int(ts_ns[i])
        if dt_vote > dt_kf:                                                                  # :931-937
            cur[27] = 1.0
            kfs.append(dict(T_f_w=T.copy(), T_prior=T.copy(), imu=cur, ts=int(ts_ns[i])))
            win, pairs = _kf_window(kfs, bacc_noise, bgyr_noise)
            assert optimizer.localMapVIOptimization(win, 1)
            stats.append(optimizer.last_stats)
            fr = kfs[::-1]
            for f, rec in enumerate(fr):                                                     # the write-back lands in the objects
                rec["T_f_w"] = win.T_f_w[f].copy()
                rec["imu"][15:27] = win.T_f_w[f]
                rec["imu"][12:15], rec["imu"][6:9], rec["imu"][9:12] = win.v[f], win.ba[f], win.bg[f]
            for k, (_, j) in enumerate(pairs):
                fr[j]["imu"][28:37], fr[j]["imu"][37:40], fr[j]["imu"][40:43] = win.imu_dR[k], win.imu_dv[k], win.imu_dp[k]
            tsp = ts_f[i]
        if len(kfs) == n_kf:                                                                 # :940-941
            break
    return kfs, stats


def check_euroc_bias(kfs):
    """Assertions of imu_test.cpp:944-945 on the 30th keyframe."""
    assert len(kfs) == 30
    assert np.linalg.norm(orc.imu_get(kfs[29]["imu"], "ba") - EUROC_BA) < 0.02
    assert np.linalg.norm(orc.imu_get(kfs[29]["imu"], "bg") - EUROC_BG) < 0.02


# ----------------------------------------------------------------------------------------------------------------
# imu_test.cpp:813-880 — the initialisation part of simuEuroc: 10 keyframes 0.5 s apart on the EuRoC ground truth from sample
# 2000 on, IMU samples differentiated from it (no bias), poses and velocities scaled by 0.5; VIInit(local_map, R_w_i, true)
# must bring every keyframe pose back onto the ground truth (0.02).
# ----------------------------------------------------------------------------------------------------------------
def euroc_viinit_window(n_kf=10, dt_kf=0.5, scale_factor=0.5):
    """Returns (window, ground-truth T_w_f 4x4 per frame of the window, newest first)."""
    acc, gyr, R, p, v, ts_f, ts_ns = _euroc_samples()
    rate = 200.0
    eta6 = eta(rate)                                                                        # SetUp(), :60-66
    T0 = _T_f_w(R[0], scale_factor * p[0])                                                   # :826-828
    imu0 = orc.imu_state(acc[0], gyr[0], T_f_w=T0, v=scale_factor * v[0], is_kf=True)        # :823-831
    kfs = [dict(T_f_w=T0.copy(), imu=imu0, ts=int(ts_ns[0]), k=0)]
    tsp, last, last_ts = ts_f[0], imu0, int(ts_ns[0])
    for i in range(1, acc.shape[0]):
        dt_vote = (ts_f[i] - tsp) * 1e-9                                                     # :838
        kf_imu = kfs[-1]["imu"]
        cur = orc.process_imu(last, orc.imu_get(kf_imu, "ba"), orc.imu_get(kf_imu, "bg"), (int(ts_ns[i]) - last_ts) * 1e-9, eta6, rate,
                              acc[i], gyr[i])                                                # :841-846
        T = _T_f_w(R[i], scale_factor * p[i])                                                # :849-852
        cur[15:27] = T
        cur[12:15] = scale_factor * v[i]                                                     # :853
        last, last_ts = cur, int(ts_ns[i])
        if dt_vote > dt_kf:                                                                  # :856-861
            cur[27] = 1.0
            kfs.append(dict(T_f_w=T.copy(), imu=cur, ts=int(ts_ns[i]), k=i))
            tsp = ts_f[i]
        if len(kfs) == n_kf:                                                                 # :864-865
            break
    for rec in kfs:
        rec["T_prior"] = rec["T_f_w"]
    win, pairs = _kf_window(kfs, BACC_NOISE, BGYR_NOISE)
    win.n_fixed = 0
    win.has_prior = np.zeros(win.n_frames, np.uint8)                                         # no setPrior in this part of the test
    # VIInit pairs every frame with its getLastKF() — no dt test (AOptimizer.cpp:485-502); at 0.5 s spacing _kf_window kept them all
    assert len(pairs) == len(kfs) - 1
    gt = []
    for rec in kfs[::-1]:
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R[rec["k"]], p[rec["k"]]
        gt.append(T)
    return win, gt


def check_euroc_viinit(win, gt, tol=0.02):
    """Assertions of imu_test.cpp:873-878 on the updated window."""
    for f in range(win.n_frames):
        T_f_w = np.vstack([win.T_f_w[f].reshape(3, 4), [0, 0, 0, 1]])
        assert np.linalg.norm(gt[f] @ T_f_w - np.eye(4)) < tol, (f, np.linalg.norm(gt[f] @ T_f_w - np.eye(4)))

"""Synthetic sliding-window generator (numpy, host side) for the BASELINE.json configurations.

Produces flattened windows (``abi.Window``) of the shapes SURVEY.md §8(d) fixes:

    C1  10 KF x  1 250 landmarks x  10 000 obs   EuRoC ground truth + the eth.yaml stereo rig (plumbing, CPU only)
    C2  20 KF x  2 000 landmarks x  15 000 obs   stereo + IMU
    C3  50 KF x 10 000 landmarks x  80 000 obs   full VIO factor set          (headline)
    C4  30 KF x  4 000 landmarks x  24 000 obs   two non-overlapping cameras, no IMU (localMapBA)
    C5 200 KF x 100 000 landmarks x 800 000 obs  C3 pattern scaled

Everything the reference front end would have produced is synthesised here: ground-truth trajectory, 200 Hz IMU
samples, pre-integrated deltas / covariance / bias Jacobians (a numpy restatement of ``IMU::processIMU``,
reference cpp/src/data/sensors/IMU.cpp:5-91, including its quirks), stereo observations with 1 px noise and their
bearing vectors (``Camera::getRayCamera``, Camera.cpp:15-25), and a perturbed initial state.

This module never imports ``oracle``; tests compare its pre-integration against the oracle's restatement.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from . import abi

GRAVITY = np.array([0.0, 0.0, -9.81])  # data/sensors/IMU.h:8

# EuRoC rig, ros/config/dataset/eth.yaml:5-39 (T_BS = sensor->body; the optimizer uses T_s_f = T_BS^-1).
T_BS_CAM0 = np.array([
    [0.0148655429818, -0.999880929698, 0.00414029679422, -0.0216401454975],
    [0.999557249008, 0.0149672133247, 0.025715529948, -0.064676986768],
    [-0.0257744366974, 0.00375618835797, 0.999660727178, 0.00981073058949],
    [0.0, 0.0, 0.0, 1.0]])
T_BS_CAM1 = np.array([
    [0.0125552670891, -0.999755099723, 0.0182237714554, -0.0198435579556],
    [0.999598781151, 0.0130119051815, 0.0251588363115, 0.0453689425024],
    [-0.0253898008918, 0.0179005838253, 0.999517347078, 0.00786212447038],
    [0.0, 0.0, 0.0, 1.0]])
K_CAM0 = np.array([458.654, 457.296, 367.215, 248.375])
K_CAM1 = np.array([457.587, 456.134, 379.999, 255.238])

# Non-overlapping rig, cpp/tests/nofov_test.cpp:78-82 (T_f_s = sensor->frame).
T_F_S1 = np.array([
    [-0.01404322, 0.00230685, 0.99989873, 0.06684756],
    [-0.99986816, -0.00818516, -0.01402391, 0.23005136],
    [0.00815198, -0.99996384, 0.00242149, 0.01394674],
    [0.0, 0.0, 0.0, 1.0]])
T_F_S2 = np.array([
    [0.0279097, 0.00437207, -0.99960089, -0.06755216],
    [0.99961045, -0.00016776, 0.02790923, -0.2074177],
    [-0.00004568, -0.99999043, -0.00437504, 0.0111566],
    [0.0, 0.0, 0.0, 1.0]])

# IMU noise model, cpp/tests/imu_test.cpp:64-68
GYR_NOISE = (0.5 * np.pi) / (180 * 60)
BGYR_NOISE = 1.9393e-05
ACC_NOISE = 0.1 / 60
BACC_NOISE = 3.0e-3
RATE_HZ = 200.0


# ------------------------------------------------------------------------------------------------------------------
# SO(3) helpers with the reference's small-angle branches (include/utilities/geometry.h:17-37,131-166)
# ------------------------------------------------------------------------------------------------------------------
def skew(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)


def exp_so3(v):
    v = np.asarray(v, dtype=np.float64)
    angle = np.linalg.norm(v)
    if angle < 1e-9:
        return np.eye(3) + skew(v)
    k = skew(v / angle)
    return np.eye(3) + (1.0 - np.cos(angle)) * (k @ k) + np.sin(angle) * k


def log_so3(M):
    c = min(max(0.5 * np.trace(M) - 0.5, -1.0), 1.0)
    angle = np.arccos(c)
    A = M - M.T
    w = np.array([A[2, 1], A[0, 2], A[1, 0]])
    if abs(np.sin(angle)) < 1e-9 or angle < 1e-9:
        return 0.5 * w
    return (0.5 * angle / np.sin(angle)) * w


def right_jacobian(w):
    w = np.asarray(w, dtype=np.float64)
    n = np.linalg.norm(w)
    if n < 1e-5:
        return np.eye(3)
    S = skew(w)
    return np.eye(3) - ((1 - np.cos(n)) / (n * n)) * S + ((n - np.sin(n)) / (n ** 3)) * (S @ S)


def inv_T(T):
    """Inverse of a 4x4 (or 3x4) rigid transform, returned as 4x4."""
    R, t = T[:3, :3], T[:3, 3]
    out = np.eye(4)
    out[:3, :3] = R.T
    out[:3, 3] = -R.T @ t
    return out


def T34(T):
    return np.ascontiguousarray(T[:3, :4]).reshape(12)


# ------------------------------------------------------------------------------------------------------------------
# IMU pre-integration (numpy restatement of IMU::processIMU, IMU.cpp:5-91)
# ------------------------------------------------------------------------------------------------------------------
@dataclass
class Preint:
    dR: np.ndarray
    dv: np.ndarray
    dp: np.ndarray
    cov: np.ndarray
    J_dR_bg: np.ndarray
    J_dv_ba: np.ndarray
    J_dv_bg: np.ndarray
    J_dp_ba: np.ndarray
    J_dp_bg: np.ndarray


def preintegrate(acc, gyr, dt, ba, bg, eta, prev: Preint | None = None) -> Preint:
    """Pre-integrate samples acc[k], gyr[k] (k = 0..n-1, sample 0 belongs to the keyframe) over n steps of `dt`.

    All samples of the interval carry the keyframe's bias estimate (IMU.cpp:17-18).  `prev` is the keyframe IMU's own
    pre-integration from the interval before: the reference reads ``_last_IMU->getDeltaR()`` for the noise matrix B even
    on the first step after a keyframe (IMU.cpp:44-47), so that stale delta_R leaks into Sigma.
    """
    n = acc.shape[0]
    dt22 = 0.5 * dt * dt
    Eta = np.diag(eta)
    last_dR = np.eye(3) if prev is None else prev.dR
    st = None
    for k in range(n):
        a = acc[k] - ba
        dv = a * dt
        dp = a * dt22
        dR = exp_so3((gyr[k] - bg) * dt)
        Jrk = right_jacobian((gyr[k] - bg) * dt)
        B = np.zeros((9, 6))
        B[0:3, 0:3] = Jrk * dt
        B[3:6, 3:6] = last_dR * dt
        B[6:9, 3:6] = last_dR * dt22
        if k == 0:
            Sigma = B @ Eta @ B.T
            Sigma[6:9, 6:9] += 0.0001 * np.eye(3) * dt
            st = Preint(dR, dv, dp, Sigma, -Jrk * dt, -np.eye(3) * dt, np.zeros((3, 3)), -dt22 * np.eye(3), np.zeros((3, 3)))
        else:
            dR_dA = st.dR @ skew(a)
            A = np.eye(9)
            A[0:3, 0:3] = dR.T
            A[3:6, 0:3] = -dR_dA * dt
            A[6:9, 0:3] = -dR_dA * dt22
            A[6:9, 3:6] = np.eye(3) * dt
            Sigma = A @ st.cov @ A.T + B @ Eta @ B.T
            Sigma[6:9, 6:9] += 0.0001 * np.eye(3) * dt
            st = Preint(
                st.dR @ dR,
                st.dv + st.dR @ dv,
                st.dp + st.dv * dt + st.dR @ dp,
                Sigma,
                dR.T @ st.J_dR_bg - Jrk * dt,
                st.J_dv_ba - st.dR * dt,
                st.J_dv_bg - dR_dA @ st.J_dR_bg * dt,
                st.J_dp_ba + st.J_dv_ba * dt - dt22 * st.dR,
                st.J_dp_bg + st.J_dv_bg * dt - dt22 * (dR_dA @ st.J_dR_bg),
            )
        last_dR = st.dR
    return st


# ------------------------------------------------------------------------------------------------------------------
# camera helpers
# ------------------------------------------------------------------------------------------------------------------
def project(K, T_s_w, p_w):
    """Pinhole projection of world points p_w [N,3] through T_s_w (4x4). Returns uv [N,2], z [N]."""
    pc = p_w @ T_s_w[:3, :3].T + T_s_w[:3, 3]
    z = pc[:, 2]
    uv = np.stack([K[0] * pc[:, 0] / z + K[2], K[1] * pc[:, 1] / z + K[3]], axis=1)
    return uv, z


def ray_camera(K, uv):
    """Camera::getRayCamera (Camera.cpp:15-25)."""
    r = np.stack([(uv[:, 0] - K[2]) / K[0], (uv[:, 1] - K[3]) / K[1], np.ones(uv.shape[0])], axis=1)
    return r / np.linalg.norm(r, axis=1, keepdims=True)


# ------------------------------------------------------------------------------------------------------------------
# configurations
# ------------------------------------------------------------------------------------------------------------------
@dataclass
class SynthConfig:
    name: str
    n_frames: int
    n_lmks: int
    span: int = 4                  # keyframes each landmark is seen from
    short_every: int = 0           # every k-th landmark is seen from only 2 keyframes (C2)
    vio: bool = True
    stereo: bool = True            # both cameras see every landmark; False: non-overlapping rig, one camera per landmark
    factor_kind: int = abi.SDV_FACTOR_ANGULAR
    n_fixed: int = 1               # ros/config/config.yaml:35
    kf_dt: float = 0.25
    pixel_noise: float = 1.0
    seed: int = 20260925
    perturb: bool = True
    prior_on_oldest: bool = True
    trajectory: str = "circle"     # "circle": the analytic 6-dof curve of SURVEY.md §8d; "euroc": the EuRoC ground truth (C1)


CONFIGS = {
    # BASELINE config 1 (plumbing, CPU only): 10 keyframes 0.5 s apart on the EuRoC ground truth from sample 2000 on, IMU by
    # differentiating it as cpp/tests/imu_test.cpp:741-757 does, the stereo rig of eth.yaml, landmarks sampled in view.
    "C1": SynthConfig("C1", 10, 1250, span=4, kf_dt=0.5, seed=20260925 + 1, trajectory="euroc"),
    "tiny": SynthConfig("tiny", 6, 40, span=3, seed=20260925 + 100),
    "small": SynthConfig("small", 10, 300, span=4, seed=20260925 + 101),
    "C2": SynthConfig("C2", 20, 2000, span=4, short_every=8, seed=20260925 + 2),
    "C3": SynthConfig("C3", 50, 10000, span=4, seed=20260925 + 3),
    "C4": SynthConfig("C4", 30, 4000, span=6, vio=False, stereo=False, seed=20260925 + 4),
    "C5": SynthConfig("C5", 200, 100000, span=4, seed=20260925 + 5),
}


def _integrate(R, acc_clean, p0, v0, dt):
    """Position / velocity by the recursion of IMU.cpp:35-41 with the true (noise-free, bias-free) specific force."""
    n = R.shape[0]
    p = np.zeros((n, 3))
    v = np.zeros((n, 3))
    p[0], v[0] = p0, v0
    for k in range(n - 1):
        a_b = acc_clean[k]
        v[k + 1] = v[k] + R[k] @ a_b * dt + GRAVITY * dt
        p[k + 1] = p[k] + v[k] * dt + R[k] @ a_b * (0.5 * dt * dt) + GRAVITY * (0.5 * dt * dt)
    return p, v


def _circle_trajectory(n_samples, dt):
    """Smooth 6-dof curve: circle r = 3 m at 0.5 m/s with +-10 deg roll / pitch oscillation."""
    t = np.arange(n_samples) * dt
    radius, speed = 3.0, 0.5
    yaw_rate = speed / radius
    R0 = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])  # body z (optical axis) -> world x
    amp = np.deg2rad(10.0)

    def R_of(tt):
        psi = yaw_rate * tt
        Rz = np.array([[np.cos(psi), -np.sin(psi), 0], [np.sin(psi), np.cos(psi), 0], [0, 0, 1]])
        osc = exp_so3(np.array([amp * np.sin(0.9 * tt), amp * np.sin(0.7 * tt + 0.5), 0.0]))
        return Rz @ R0 @ osc

    def acc_world(tt):
        psi = yaw_rate * tt
        a = -radius * yaw_rate ** 2 * np.array([np.cos(psi), np.sin(psi), 0.0])
        a[2] = -0.2 * (0.8 ** 2) * np.sin(0.8 * tt)
        return a

    R = np.stack([R_of(tt) for tt in t])
    gyr_clean = np.zeros((n_samples, 3))
    acc_clean = np.zeros((n_samples, 3))
    for k in range(n_samples):
        if k + 1 < n_samples:
            gyr_clean[k] = log_so3(R[k].T @ R[k + 1]) / dt
        else:
            gyr_clean[k] = gyr_clean[k - 1]
        acc_clean[k] = R[k].T @ (acc_world(t[k]) - GRAVITY)
    p, v = _integrate(R, acc_clean, np.array([radius, 0.0, 0.0]), np.array([0.0, speed, 0.2 * 0.8]), dt)
    return R, p, v, acc_clean, gyr_clean


EUROC_SLICE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "euroc_gt_slice.npz")


def load_euroc_slice():
    """The committed rows of the reference's cpp/tests/euroc_gt.csv (tests/golden/make_euroc_slice.py): timestamps [s],
    rotations body -> world, positions, velocities.  Rotations are the polar factors of Quaterniond(w,x,y,z).toRotationMatrix(),
    which is what Eigen's Affine3d::rotation() returns for the 6-digit quaternions of the file (imu_test.cpp:21-27)."""
    z = np.load(EUROC_SLICE)
    q = z["q_wxyz"]
    R = np.zeros((q.shape[0], 3, 3))
    for k, (w, x, y, zz) in enumerate(q):
        M = np.array([[1 - 2 * (y * y + zz * zz), 2 * (x * y - w * zz), 2 * (x * zz + w * y)],
                      [2 * (x * y + w * zz), 1 - 2 * (x * x + zz * zz), 2 * (y * zz - w * x)],
                      [2 * (x * zz - w * y), 2 * (y * zz + w * x), 1 - 2 * (x * x + y * y)]])
        U, _, Vt = np.linalg.svd(M)
        R[k] = U @ Vt
    return z["timestamp_ns"].astype(np.float64) * 1e-9, R, z["p"].copy(), z["v"].copy()


def _euroc_trajectory(n_samples, dt):
    """EuRoC ground truth from sample 2000 on, measurements by differentiating it exactly as the reference's simuEuroc does
    (imu_test.cpp:741-757: acc from the velocity difference rotated by the CURRENT attitude, gyr from the attitude
    increment); position / velocity re-integrated from them so that the IMU factors and the poses agree."""
    ts, R, p_gt, v_gt = load_euroc_slice()
    assert n_samples + 1 <= ts.size, "the committed EuRoC slice is too short for this window"
    acc_clean = np.zeros((n_samples, 3))
    gyr_clean = np.zeros((n_samples, 3))
    for k in range(n_samples):
        h = ts[k + 1] - ts[k]
        acc_clean[k] = (1 / h) * (R[k + 1].T @ (v_gt[k + 1] - v_gt[k])) - R[k + 1].T @ GRAVITY
        gyr_clean[k] = (1 / h) * log_so3(R[k].T @ R[k + 1])
    R = R[:n_samples].copy()
    p, v = _integrate(R, acc_clean, p_gt[0], v_gt[0], dt)
    return R, p, v, acc_clean, gyr_clean


def make_window(cfg: SynthConfig | str, **overrides) -> abi.Window:
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    if overrides:
        cfg = SynthConfig(**{**cfg.__dict__, **overrides})
    rng = np.random.Generator(np.random.MT19937(cfg.seed))
    F = cfg.n_frames
    steps = int(round(cfg.kf_dt * RATE_HZ))
    dt = 1.0 / RATE_HZ
    n_samples = (F - 1) * steps + 1

    # ---- ground-truth orientation and IMU-consistent position/velocity (integrated as the reference does)
    R, p, v, acc_clean, gyr_clean = (_euroc_trajectory if cfg.trajectory == "euroc" else _circle_trajectory)(n_samples, dt)
    ba_true = np.array([0.02, -0.01, 0.015])
    bg_true = np.array([0.001, -0.002, 0.0015])

    kf_idx = np.arange(F) * steps               # time order: keyframe k at sample kf_idx[k]
    T_w_f_gt = np.zeros((F, 4, 4))
    for k in range(F):
        T_w_f_gt[k] = np.eye(4)
        T_w_f_gt[k, :3, :3] = R[kf_idx[k]]
        T_w_f_gt[k, :3, 3] = p[kf_idx[k]]
    T_f_w_gt = np.stack([inv_T(T) for T in T_w_f_gt])
    v_gt = v[kf_idx]

    # ---- cameras
    if cfg.stereo:
        T_s_f = np.stack([inv_T(T_BS_CAM0), inv_T(T_BS_CAM1)])
        K = np.stack([K_CAM0, K_CAM1])
    else:
        T_s_f = np.stack([inv_T(T_F_S1), inv_T(T_F_S2)])
        K = np.stack([K_CAM0, K_CAM0])
    n_cams = 2

    # ---- landmarks + observations (time-ordered, like a SLAM front end creates them); fully vectorised
    L = cfg.n_lmks
    margin = 20.0
    spans = np.full(L, cfg.span)
    if cfg.short_every:
        spans[np.arange(L) % cfg.short_every == cfg.short_every - 1] = 2
    k0 = np.minimum((np.arange(L) * (F - cfg.span + 1)) // L, F - spans)
    lmk_cam = (np.arange(L) % 2) if not cfg.stereo else np.zeros(L, dtype=np.int64)
    ncl = 2 if cfg.stereo else 1
    T_s_f44 = np.stack([np.vstack([T[:3], [0, 0, 0, 1]]) for T in T_s_f])
    T_s_w = np.einsum("cij,kjl->kcil", T_s_f44, T_f_w_gt)          # [F, C, 4, 4]
    counts = spans * ncl
    ptr = np.concatenate([[0], np.cumsum(counts)])
    n_obs = int(ptr[-1])
    obs_lmk = np.repeat(np.arange(L), counts)
    within = np.arange(n_obs) - ptr[obs_lmk]
    obs_k = k0[obs_lmk] + within // ncl                              # time index of the observing keyframe
    obs_cam = (within % ncl) if cfg.stereo else lmk_cam[obs_lmk]

    def project_obs(sel_obs, pts):
        Tm = T_s_w[obs_k[sel_obs], obs_cam[sel_obs]]
        pc = np.einsum("nij,nj->ni", Tm[:, :3, :3], pts) + Tm[:, :3, 3]
        Ko = K[obs_cam[sel_obs]]
        z = pc[:, 2]
        uv = np.stack([Ko[:, 0] * pc[:, 0] / z + Ko[:, 2], Ko[:, 1] * pc[:, 1] / z + Ko[:, 3]], axis=1)
        inb = (z > 0.5) & (uv[:, 0] >= margin) & (uv[:, 0] <= 2 * Ko[:, 2] - margin) & \
              (uv[:, 1] >= margin) & (uv[:, 1] <= 2 * Ko[:, 3] - margin)
        return uv, inb

    lmk_gt = np.zeros((L, 3))
    todo = np.arange(L)
    for _ in range(200):
        if not todo.size:
            break
        n = todo.size
        c = lmk_cam[todo]
        kc = k0[todo] + np.minimum(1, spans[todo] - 1)
        Kc = K[c]
        u = rng.uniform(margin, 2 * Kc[:, 2] - margin, n)
        vv = rng.uniform(margin, 2 * Kc[:, 3] - margin, n)
        depth = rng.uniform(2.0, 15.0, n)
        pc = np.stack([(u - Kc[:, 2]) / Kc[:, 0] * depth, (vv - Kc[:, 3]) / Kc[:, 1] * depth, depth], axis=1)
        Tm = T_s_w[kc, c]
        Rm, tm = Tm[:, :3, :3], Tm[:, :3, 3]
        pw = np.einsum("nji,nj->ni", Rm, pc - tm)                   # inverse transform
        lmk_gt[todo] = pw
        mask = np.zeros(L, dtype=bool)
        mask[todo] = True
        sel = np.nonzero(mask[obs_lmk])[0]
        _, inb = project_obs(sel, lmk_gt[obs_lmk[sel]])
        bad = np.zeros(L, dtype=np.int64)
        np.add.at(bad, obs_lmk[sel], (~inb).astype(np.int64))
        todo = todo[bad[todo] > 0]
    assert not todo.size, "landmark sampling did not converge"
    uv_exact, inb = project_obs(np.arange(n_obs), lmk_gt[obs_lmk])
    assert inb.all()
    obs_frame = (F - 1 - obs_k).astype(np.int32)                    # frames are stored NEWEST -> OLDEST (amap.h:28-32)
    obs_lmk = obs_lmk.astype(np.int32)
    obs_cam = obs_cam.astype(np.int32)
    obs_uv = uv_exact + rng.normal(0.0, cfg.pixel_noise, (n_obs, 2))
    obs_bearing = np.zeros((n_obs, 3))
    for cc in range(n_cams):
        m = obs_cam == cc
        obs_bearing[m] = ray_camera(K[cc], obs_uv[m])

    # ---- IMU measurements, bias estimates, pre-integration
    win_kwargs = {}
    if cfg.vio:
        gyr_meas = gyr_clean + bg_true + rng.normal(0.0, GYR_NOISE * np.sqrt(RATE_HZ), (n_samples, 3))
        acc_meas = acc_clean + ba_true + rng.normal(0.0, ACC_NOISE * np.sqrt(RATE_HZ), (n_samples, 3))
        if cfg.perturb:
            ba_est = ba_true + rng.normal(0, 0.02, 3) + rng.normal(0, 1e-3, (F, 3))
            bg_est = bg_true + rng.normal(0, 0.002, 3) + rng.normal(0, 1e-4, (F, 3))
            # the fixed (oldest) keyframes are the converged part of the window: their states stay at ground truth
            ba_est[:cfg.n_fixed] = ba_true
            bg_est[:cfg.n_fixed] = bg_true
        else:
            ba_est = np.tile(ba_true, (F, 1))
            bg_est = np.tile(bg_true, (F, 1))
        eta = np.array([GYR_NOISE ** 2] * 3 + [ACC_NOISE ** 2] * 3) * RATE_HZ  # IMU.h:39-41
        P = F - 1
        pre = []
        prev = None
        for k in range(P):  # interval keyframe k -> k+1 (time order)
            a, b = kf_idx[k], kf_idx[k + 1]
            prev = preintegrate(acc_meas[a:b], gyr_meas[a:b], dt, ba_est[k], bg_est[k], eta, prev)
            pre.append(prev)
        # imu entries follow frame_vector order of j (newest first): j index = F-1-(k+1), i index = F-1-k
        order = list(range(P - 1, -1, -1))
        win_kwargs.update(
            imu_i=np.array([F - 1 - k for k in order], dtype=np.int32),
            imu_j=np.array([F - 1 - (k + 1) for k in order], dtype=np.int32),
            imu_dt=np.full(P, cfg.kf_dt),
            imu_dR=np.stack([pre[k].dR.reshape(9) for k in order]),
            imu_dv=np.stack([pre[k].dv for k in order]),
            imu_dp=np.stack([pre[k].dp for k in order]),
            imu_cov=np.stack([pre[k].cov.reshape(81) for k in order]),
            imu_J_dR_bg=np.stack([pre[k].J_dR_bg.reshape(9) for k in order]),
            imu_J_dv_ba=np.stack([pre[k].J_dv_ba.reshape(9) for k in order]),
            imu_J_dv_bg=np.stack([pre[k].J_dv_bg.reshape(9) for k in order]),
            imu_J_dp_ba=np.stack([pre[k].J_dp_ba.reshape(9) for k in order]),
            imu_J_dp_bg=np.stack([pre[k].J_dp_bg.reshape(9) for k in order]),
            imu_sigma_ba=np.full(P, BACC_NOISE),
            imu_sigma_bg=np.full(P, BGYR_NOISE),
        )

    # ---- initial (perturbed) state, flattened newest -> oldest
    rev = np.arange(F)[::-1]
    T_init = np.zeros((F, 12))
    for fi, k in enumerate(rev):
        T = T_f_w_gt[k].copy()
        if cfg.perturb and fi < F - cfg.n_fixed:
            d = np.eye(4)
            d[:3, :3] = exp_so3(rng.normal(0, 0.01, 3))
            d[:3, 3] = rng.normal(0, 0.05, 3)
            T = T @ d
        T_init[fi] = T34(T)
    lmk_init = lmk_gt + (rng.normal(0, 0.05, (L, 3)) if cfg.perturb else 0.0)
    has_prior = np.zeros(F, dtype=np.uint8)
    T_prior = np.tile(np.eye(3, 4).reshape(12), (F, 1))
    inf_prior = np.zeros((F, 6))
    if cfg.prior_on_oldest:
        has_prior[F - 1] = 1
        T_prior[F - 1] = T34(T_f_w_gt[0])
        inf_prior[F - 1] = 100.0  # imu_test.cpp:467

    if cfg.vio:
        v_init = v_gt[rev] + (rng.normal(0, 0.05, (F, 3)) if cfg.perturb else 0.0)
        if cfg.perturb and cfg.n_fixed:
            v_init[F - cfg.n_fixed:] = v_gt[rev][F - cfg.n_fixed:]
        win_kwargs.update(v=v_init, ba=ba_est[rev].copy(), bg=bg_est[rev].copy(), has_imu=np.ones(F, dtype=np.uint8))

    win = abi.Window(
        vio=cfg.vio, factor_kind=cfg.factor_kind, n_fixed=cfg.n_fixed,
        T_f_w=T_init, T_s_f=np.stack([T34(T.reshape(4, 4)) for T in T_s_f]), K=K.copy(), lmk_t=lmk_init,
        obs_lmk=obs_lmk, obs_frame=obs_frame, obs_cam=obs_cam, obs_bearing=obs_bearing, obs_uv=obs_uv,
        has_prior=has_prior, T_prior=T_prior, inf_prior=inf_prior, **win_kwargs,
    )
    win.meta = dict(
        name=cfg.name, T_f_w_gt=np.stack([T34(T_f_w_gt[k]) for k in rev]), lmk_gt=lmk_gt,
        v_gt=v_gt[rev] if cfg.vio else None, ba_true=ba_true, bg_true=bg_true, cfg=cfg,
    )
    return win.normalise()


def add_dense_prior(win: abi.Window, n_keep: int = 20, seed: int = 1, with_frame: bool = True) -> abi.Window:
    """Attach a synthetic dense marginalisation prior (isae::MarginalizationFactor, marginalization.hpp:88-218):
    r = r0 + J dx over the oldest free keyframe's (pose, v, ba, bg) and `n_keep` landmarks that keyframe observes.
    J = Lambda^1/2 U^T of a random well-conditioned information matrix, as computeJacobiansAndResiduals produces
    (marginalization.cpp:516-530); a few landmarks carry keep_col = -1 to exercise the skip at marginalization.hpp:139."""
    rng = np.random.Generator(np.random.MT19937(seed))
    F = win.n_frames
    f_keep = F - 1 - win.n_fixed if with_frame else -1
    f_obs = F - 1 - win.n_fixed
    cand = np.unique(win.obs_lmk[win.obs_frame == f_obs])
    keep = np.sort(rng.choice(cand, size=min(n_keep, cand.size), replace=False)).astype(np.int32)
    cols, c = [], 15 if with_frame else 0
    for k in range(keep.size):
        if k % 7 == 6:
            cols.append(-1)
        else:
            cols.append(c)
            c += 3
    n = c
    sig = np.array([200.0] * 3 + [100.0] * 3 + [20.0] * 3 + [100.0] * 3 + [2000.0] * 3 if with_frame else [])
    sig = np.concatenate([sig, np.full(n - sig.size, 10.0)])
    M = rng.normal(0, 1, (n, n))
    Q, _ = np.linalg.qr(M)
    A = np.diag(sig) @ (np.eye(n) + 0.2 * Q)       # square-root information with cross terms
    J = A
    r0 = J @ rng.normal(0, 1e-3, n) * 0.5
    win.dense_prior = abi.DensePrior(J=np.ascontiguousarray(J), r0=r0, frame=f_keep, frame_col=0, keep_lmk=keep,
                                     keep_col=np.array(cols, dtype=np.int32))
    return win


def _rand_sqrt_inf(rng, n, scale, mix=0.15):
    M = rng.normal(0, 1, (n, n))
    Q, _ = np.linalg.qr(M)
    return np.diag(np.broadcast_to(scale, (n,)).astype(np.float64)) @ (np.eye(n) + mix * Q)


def add_sparse_prior_vio(win: abi.Window, n_keep: int = 20, seed: int = 2) -> abi.Window:
    """Sparsified VIO prior (AngularAdjustmentCERESAnalytic.cpp:390-424): IMUPriordx on the oldest free keyframe + one
    PoseToLandmarkFactor per kept landmark (relative position of the landmark in that keyframe)."""
    rng = np.random.Generator(np.random.MT19937(seed))
    F = win.n_frames
    f = F - 1 - win.n_fixed
    gt = win.meta
    T_gt = np.vstack([gt["T_f_w_gt"][f].reshape(3, 4), [0, 0, 0, 1]])
    cand = np.unique(win.obs_lmk[win.obs_frame == f])
    keep = np.sort(rng.choice(cand, size=min(n_keep, cand.size), replace=False)).astype(np.int32)
    delta = (gt["lmk_gt"][keep] @ T_gt[:3, :3].T + T_gt[:3, 3]) + rng.normal(0, 0.01, (keep.size, 3))
    sq = np.stack([_rand_sqrt_inf(rng, 3, 30.0).reshape(9) for _ in range(keep.size)])
    scale = np.array([300.0] * 3 + [150.0] * 3 + [30.0] * 3 + [200.0] * 3 + [3000.0] * 3)
    win.sparse_prior = abi.SparsePrior(
        has_imu_prior=True, frame=f, T_prior=gt["T_f_w_gt"][f].copy(), v_prior=gt["v_gt"][f].copy(), ba_prior=gt["ba_true"].copy(),
        bg_prior=gt["bg_true"].copy(), imu_sqrt_inf=_rand_sqrt_inf(rng, 15, scale).reshape(225), p2l_lmk=keep, p2l_delta=delta, p2l_sqrt_inf=sq)
    return win


def add_sparse_prior_vo(win: abi.Window, n_keep: int = 100, seed: int = 3) -> abi.Window:
    """Sparsified VO prior (AngularAdjustmentCERESAnalytic.cpp:426-482): Landmark3DPrior on the first kept landmark and a
    chain of LandmarkToLandmarkFactor between consecutive kept landmarks (the Chow-Liu chain of sparsifyVO)."""
    rng = np.random.Generator(np.random.MT19937(seed))
    F = win.n_frames
    f = F - 1 - win.n_fixed
    gt = win.meta
    cand = np.unique(win.obs_lmk[win.obs_frame == f])
    keep = rng.choice(cand, size=min(n_keep, cand.size), replace=False).astype(np.int32)  # chain order, not sorted
    a, b = keep[:-1].copy(), keep[1:].copy()
    delta = gt["lmk_gt"][a] - gt["lmk_gt"][b] + rng.normal(0, 0.01, (a.size, 3))
    win.sparse_prior = abi.SparsePrior(
        has_lmk_prior=True, lmk0=int(keep[0]), lmk_prior=gt["lmk_gt"][keep[0]] + rng.normal(0, 0.01, 3),
        lmk_sqrt_inf=_rand_sqrt_inf(rng, 3, 20.0).reshape(9), l2l_a=a, l2l_b=b, l2l_delta=delta,
        l2l_sqrt_inf=np.stack([_rand_sqrt_inf(rng, 3, 40.0).reshape(9) for _ in range(a.size)]))
    return win


def make_c4() -> abi.Window:
    """BASELINE config 4: 30 KF, two non-overlapping cameras, no IMU, sparsified VO marginal prior (unary + 99-link chain)."""
    return add_sparse_prior_vo(make_window("C4"), n_keep=100)


def apply_delta(win: abi.Window, d: abi.Delta) -> dict:
    """State write-back of AOptimizer.cpp:391-418 on numpy copies: returns the updated state arrays."""
    F = win.n_frames
    T_new = np.zeros((F, 12))
    for f in range(F):
        T = np.vstack([win.T_f_w[f].reshape(3, 4), [0, 0, 0, 1]])
        dT = np.eye(4)
        dT[:3, :3] = exp_so3(d.dpose[f, :3])
        dT[:3, 3] = d.dpose[f, 3:]
        T_new[f] = T34(T @ dT)
    out = dict(T_f_w=T_new, lmk_t=win.lmk_t + d.dlmk)
    if win.vio:
        out.update(v=win.v + d.dv, ba=win.ba + d.dba, bg=win.bg + d.dbg)
    return out


IMU_FACTOR_FIELDS = ("imu_i", "imu_j", "imu_dt", "imu_dR", "imu_dv", "imu_dp", "imu_cov", "imu_J_dR_bg", "imu_J_dv_ba", "imu_J_dv_bg",
                     "imu_J_dp_ba", "imu_J_dp_bg", "imu_sigma_ba", "imu_sigma_bg")


def skip_imu_factor(win: abi.Window, frame_j: int) -> abi.Window:
    """Models a keyframe more than 1 s after its previous keyframe: the IMUFactor / IMUBiasFactor of `frame_j` disappear
    from the window (AOptimizer.cpp:69) while its pre-integration stays on the host side for the write-back, which corrects
    it regardless of the gap (AOptimizer.cpp:421-434)."""
    win.normalise()
    ps = [p for p in range(win.n_imu) if int(win.imu_j[p]) == frame_j]
    assert len(ps) == 1, ps
    p = ps[0]
    old = win.skipped_preint
    cat = lambda a, b: b.copy() if a is None else np.concatenate([a, b])
    win.skipped_preint = abi.SkippedPreint(
        frame=cat(old and old.frame, win.imu_j[p:p + 1]), prev=cat(old and old.prev, win.imu_i[p:p + 1]),
        dR=cat(old and old.dR, win.imu_dR[p:p + 1]), dv=cat(old and old.dv, win.imu_dv[p:p + 1]), dp=cat(old and old.dp, win.imu_dp[p:p + 1]),
        J_dR_bg=cat(old and old.J_dR_bg, win.imu_J_dR_bg[p:p + 1]), J_dv_ba=cat(old and old.J_dv_ba, win.imu_J_dv_ba[p:p + 1]),
        J_dv_bg=cat(old and old.J_dv_bg, win.imu_J_dv_bg[p:p + 1]), J_dp_ba=cat(old and old.J_dp_ba, win.imu_J_dp_ba[p:p + 1]),
        J_dp_bg=cat(old and old.J_dp_bg, win.imu_J_dp_bg[p:p + 1]))
    keep = np.array([q for q in range(win.n_imu) if q != p], dtype=np.int64)
    for name in IMU_FACTOR_FIELDS:
        setattr(win, name, np.ascontiguousarray(getattr(win, name)[keep]))
    return win

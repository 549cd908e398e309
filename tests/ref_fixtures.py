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

"""Pin the oracle against the reference's own known-answer tests.

Every test names the reference test it re-states (cpp/tests/*.cpp).  The reference draws its random fixtures without a
seed (residual_test.cpp:18-19), so inputs are regenerated here with a fixed seed and the reference's *criterion* is kept
(and tightened: the reference compares the SUM of Jacobian differences to 1e-5, we compare every entry).
"""
import numpy as np
import pytest

from oracle import oracle as orc

RNG = np.random.default_rng(20260925)

# imu_test.cpp:64-68
GYR_NOISE = (0.5 * np.pi) / (180 * 60)
BGYR_NOISE = 1.9393e-05
ACC_NOISE = 0.1 / 60
BACC_NOISE = 3.0e-3


def eta(rate):
    e = np.array([GYR_NOISE] * 3 + [ACC_NOISE] * 3) ** 2
    return e * rate  # IMU.h:39-41


def rand_rot(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def T34(R, t):
    return np.hstack([R, np.asarray(t).reshape(3, 1)]).reshape(12)


def num_jac(f, x0, eps=1e-6):
    x0 = np.asarray(x0, dtype=float)
    f0 = f(x0)
    J = np.zeros((f0.size, x0.size))
    for i in range(x0.size):
        d = np.zeros_like(x0)
        d[i] = eps
        J[:, i] = (f(x0 + d) - f(x0 - d)) / (2 * eps)
    return J


# ----------------------------------------------------------------------------------------------------------------
# geometry.h
# ----------------------------------------------------------------------------------------------------------------
def test_so3_exp_log_roundtrip_and_branches():
    for _ in range(50):
        w = RNG.normal(size=3)
        w *= RNG.uniform(0, 3.0) / np.linalg.norm(w)  # |w| < pi
        R = orc.exp_so3(w)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-13)
        assert np.allclose(orc.log_so3(R), w, atol=1e-10)
    # first-order branch below 1e-9 (geometry.h:136-138): exactly I + skew
    w = np.array([1e-10, -2e-10, 3e-10])
    S = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    assert np.array_equal(orc.exp_so3(w), np.eye(3) + S)
    # right Jacobian snaps to identity below 1e-5 (geometry.h:33-34)
    assert np.array_equal(orc.right_jacobian(np.array([5e-6, 0, 0])), np.eye(3))
    assert not np.array_equal(orc.right_jacobian(np.array([2e-5, 0, 0])), np.eye(3))
    # Jr is the derivative of exp on the right: exp(w + d) ~ exp(w) exp(Jr d)
    w = np.array([0.3, -0.2, 0.5])
    d = 1e-6 * np.array([1.0, 2.0, -1.5])
    lhs = orc.exp_so3(w + d)
    rhs = orc.exp_so3(w) @ orc.exp_so3(orc.right_jacobian(w) @ d)
    assert np.allclose(lhs, rhs, atol=1e-11)


# ----------------------------------------------------------------------------------------------------------------
# IMU pre-integration, imu_test.cpp
# ----------------------------------------------------------------------------------------------------------------
def fixture_imu():
    """ImuTest::SetUp (imu_test.cpp:59-96): three IMU objects 0.5 s apart, frame0 is a keyframe at identity."""
    acc = np.array([0.5, 1.0, 10.81])
    gyr = np.array([0.1, 0.3, 0.1])
    imu0 = orc.imu_state(acc, gyr, is_kf=True)
    return acc, gyr, imu0


def test_ImuTestBase():  # imu_test.cpp:103-143
    acc, gyr, imu0 = fixture_imu()
    imu1 = orc.process_imu(imu0, np.zeros(3), np.zeros(3), 0.5, eta(200), 200, acc, gyr)
    dR = orc.exp_so3(gyr / 2)
    assert (dR @ orc.imu_get(imu1, "dR").reshape(3, 3).T).trace() == 3
    assert np.linalg.norm(acc / 2 - orc.imu_get(imu1, "dv")) == 0
    assert np.linalg.norm(0.5 * acc * 0.5 * 0.5 - orc.imu_get(imu1, "dp")) == 0
    # with biases (imu_test.cpp:126-142)
    ba, bg = np.array([0.1, 0.2, 0.3]), np.array([0.2, 0.3, 0.1])
    imu0b = orc.imu_state(acc, gyr, ba=ba, bg=bg, is_kf=True)
    imu1 = orc.process_imu(imu0b, ba, bg, 0.5, eta(200), 200, acc, gyr)
    dR = orc.exp_so3((gyr - bg) * 0.5)
    assert (dR @ orc.imu_get(imu1, "dR").reshape(3, 3).T).trace() == 3
    assert np.linalg.norm((acc - ba) / 2 - orc.imu_get(imu1, "dv")) == 0
    assert np.linalg.norm(0.5 * (acc - ba) * 0.5 * 0.5 - orc.imu_get(imu1, "dp")) == 0


def test_ImuNewMeas():  # imu_test.cpp:145-159
    acc, gyr, imu0 = fixture_imu()
    imu1 = orc.process_imu(imu0, np.zeros(3), np.zeros(3), 0.5, eta(200), 200, acc, gyr)
    imu2 = orc.process_imu(imu1, np.zeros(3), np.zeros(3), 0.5, eta(200), 200, acc, gyr)
    dR1 = orc.imu_get(imu1, "dR").reshape(3, 3)
    dR = orc.exp_so3(gyr)
    assert abs((dR @ orc.imu_get(imu2, "dR").reshape(3, 3).T).trace() - 3) < 1e-14
    dv = acc * 0.5 + dR1 @ acc * 0.5
    assert np.linalg.norm(dv - orc.imu_get(imu2, "dv")) < 1e-15
    dp = 0.5 * acc * 0.5 * 0.5 + orc.imu_get(imu1, "dv") * 0.5 + 0.5 * dR1 @ acc * 0.5 * 0.5
    assert np.linalg.norm(dp - orc.imu_get(imu2, "dp")) < 1e-15


def test_checkCov():  # imu_test.cpp:161-193 (GTSAM constants, trace tolerance 1e-9)
    acc, gyr, _ = fixture_imu()
    imu0 = orc.imu_state(np.array([0.1, 0, 0]), np.array([np.pi / 100, 0, 0]), is_kf=True)
    imu1 = orc.process_imu(imu0, np.zeros(3), np.zeros(3), 0.5, eta(2), 2, acc, gyr)
    expected = np.zeros((9, 9))
    expected[np.arange(3), np.arange(3)] = 1.0577e-08
    expected[np.arange(3, 6), np.arange(3, 6)] = 1.38889e-06
    expected[np.arange(6, 9), np.arange(6, 9)] = 5.00868e-05
    for k in range(3):
        expected[3 + k, 6 + k] = expected[6 + k, 3 + k] = 3.47222e-07
    cov = orc.imu_get(imu1, "Sigma").reshape(9, 9)
    assert abs((expected - cov).trace()) < 1e-9
    assert np.allclose(cov, expected, rtol=2e-5, atol=1e-12)  # stronger than the reference's trace check


def test_checkJacobiansBiasGyr():  # imu_test.cpp:328-361
    acc, gyr, imu0 = fixture_imu()
    imu1 = orc.process_imu(imu0, np.zeros(3), np.zeros(3), 0.5, eta(200), 200, acc, gyr)
    dt = 0.5
    J_rk = orc.right_jacobian(gyr * dt)
    exp_J_dbg = -J_rk * dt
    exp_J_dv_dba = -np.eye(3) * dt
    exp_J_dp_dba = -0.5 * np.eye(3) * dt * dt
    g = lambda s, n: orc.imu_get(s, n).reshape(3, 3)
    assert abs((g(imu1, "J_dR_bg") @ np.linalg.inv(exp_J_dbg)).trace() - 3) < 1e-9
    assert abs((g(imu1, "J_dv_ba") @ np.linalg.inv(exp_J_dv_dba)).trace() - 3) < 1e-9
    assert abs(g(imu1, "J_dv_bg").sum()) < 1e-9 and abs(g(imu1, "J_dp_bg").sum()) < 1e-9
    assert abs((g(imu1, "J_dp_ba") - exp_J_dp_dba).sum()) < 1e-9
    imu2 = orc.process_imu(imu1, np.zeros(3), np.zeros(3), 0.5, eta(200), 200, acc, gyr)
    dR = orc.exp_so3(gyr * dt)
    S = np.array([[0, -acc[2], acc[1]], [acc[2], 0, -acc[0]], [-acc[1], acc[0], 0]])
    dR1 = g(imu1, "dR")
    e_dbg = dR.T @ g(imu1, "J_dR_bg") - J_rk * dt
    e_dv_dba = exp_J_dv_dba - dR1 * dt
    e_dv_dbg = -dR1 @ S @ g(imu1, "J_dR_bg") * dt
    e_dp_dba = exp_J_dp_dba + g(imu1, "J_dv_ba") * dt - 0.5 * dR1 * dt * dt
    e_dp_dbg = g(imu1, "J_dv_bg") * dt - 0.5 * dR1 @ S @ g(imu1, "J_dR_bg") * dt * dt
    assert np.allclose(g(imu2, "J_dR_bg"), e_dbg, atol=1e-12)
    assert np.allclose(g(imu2, "J_dv_ba"), e_dv_dba, atol=1e-12)
    assert np.allclose(g(imu2, "J_dv_bg"), e_dv_dbg, atol=1e-12)
    assert np.allclose(g(imu2, "J_dp_ba"), e_dp_dba, atol=1e-12)
    assert np.allclose(g(imu2, "J_dp_bg"), e_dp_dbg, atol=1e-12)


def test_TestPreInteg():  # imu_test.cpp:948-995
    a, w = 0.1, np.pi / 100.0
    acc, gyr = np.array([a, 0, 0]), np.array([w, 0, 0])
    imu0 = orc.imu_state(acc, gyr, is_kf=True)
    imu1 = orc.process_imu(imu0, np.zeros(3), np.zeros(3), 0.5, eta(200), 200, acc, gyr)
    assert (orc.imu_get(imu1, "dR").reshape(3, 3) - orc.exp_so3([w * 0.5, 0, 0])).sum() == 0
    assert np.linalg.norm(orc.imu_get(imu1, "dp") - [0.5 * a * 0.25, 0, 0]) == 0
    assert np.linalg.norm(orc.imu_get(imu1, "dv") - [0.05, 0, 0]) == 0
    imu2 = orc.process_imu(imu1, np.zeros(3), np.zeros(3), 0.5, eta(200), 200, acc, gyr)
    assert abs((orc.imu_get(imu2, "dR").reshape(3, 3) - orc.exp_so3([2.0 * 0.5 * w, 0, 0])).sum()) < 1e-6
    assert np.linalg.norm(orc.imu_get(imu2, "dp") - [0.025 + 0.5 * a * 0.25 + 0.5 * 0.1 * 0.5 * 0.5, 0, 0]) < 1e-17
    dv2 = np.array([0.05, 0, 0]) + orc.exp_so3([w * 0.5, 0, 0]) @ acc * 0.5
    assert np.linalg.norm(orc.imu_get(imu2, "dv") - dv2) < 1e-17


# The rotated free-fall fixture of predictionPositionVelocity (imu_test.cpp:363-407)
R_I_F = np.array([[0.38001193, 0.16469125, 0.91020202], [0.03067918, -0.9857245, 0.16554758], [0.92447267, -0.0349858, -0.37963966]])


def free_fall():
    # The 8-digit matrix is not exactly orthonormal; the reference reads it through Eigen's Affine3d::rotation()
    # (polar decomposition) both for the accelerometer sample (imu_test.cpp:374) and inside processIMU (IMU.cpp:34),
    # so the fixture uses the polar factor.
    U, _, Vt = np.linalg.svd(R_I_F)
    T_i_f = np.eye(4)
    T_i_f[:3, :3] = U @ Vt
    T_i_f[:3, 3] = 1.0
    T_f_i = np.linalg.inv(T_i_f)
    acc = T_i_f[:3, :3].T @ np.array([0, 0, 10.81])
    gyr = np.zeros(3)
    rate = 1000.0
    kf = orc.imu_state(acc, gyr, T_f_w=T_f_i[:3].reshape(12), is_kf=True)
    last = kf
    for _ in range(1000):
        last = orc.process_imu(last, np.zeros(3), np.zeros(3), 0.001, eta(rate), rate, acc, gyr)
    return kf, last, T_i_f


def test_predictionPositionVelocity_and_IMUFactor():
    kf, cur, T_i_f = free_fall()
    T_f_w = np.vstack([orc.imu_get(cur, "T_f_w").reshape(3, 4), [0, 0, 0, 1]])
    T_w_f = np.linalg.inv(T_f_w)
    assert np.linalg.norm(T_w_f[:3, 3] - [1, 1, 1.5]) < 1e-5          # imu_test.cpp:405
    assert np.linalg.norm(orc.imu_get(cur, "v") - [0, 0, 1]) < 1e-5   # :406
    assert abs((T_w_f[:3, :3].T @ T_i_f[:3, :3]).trace() - 3) < 1e-5  # :407
    # IMUFactor residual ~ 0 at the integrated state (:440) and analytic == numeric Jacobian on all 6 blocks (:449-454)
    pre = orc.pack_preint(*[orc.imu_get(cur, n) for n in ("dR", "dv", "dp", "Sigma", "J_dR_bg", "J_dv_ba", "J_dv_bg", "J_dp_ba", "J_dp_bg")])
    T_i, T_j = orc.imu_get(kf, "T_f_w"), orc.imu_get(cur, "T_f_w")
    v_i, v_j = orc.imu_get(kf, "v"), orc.imu_get(cur, "v")
    r, J = orc.imu_factor_eval(T_i, T_j, v_i, v_j, 1.0, pre)
    assert np.linalg.norm(r) < 1e-3
    f = lambda p: orc.imu_factor_eval(T_i, T_j, v_i, v_j, 1.0, pre, p, jac=False)[0]
    Jn = num_jac(f, np.zeros(24), 1e-6)
    scale = np.abs(J).max()
    assert np.abs(J - Jn).max() < 1e-6 * scale
    # and away from zero
    p0 = RNG.normal(0, 0.02, 24)
    _, J = orc.imu_factor_eval(T_i, T_j, v_i, v_j, 1.0, pre, p0)
    Jn = num_jac(f, p0, 1e-6)
    assert np.abs(J - Jn).max() < 1e-6 * scale


def _estimate_transform(kf, cur, dt):
    """IMU::estimateTransform (IMU.cpp:93-102) followed by the pose assignment the reference tests make with it
    (imu_test.cpp:606-607): T_f_w(cur) = dT^-1 T_f_w(kf), dT = [delta_R | delta_p + R1 v_kf dt + 1/2 R1 g dt^2]."""
    g = np.array([0.0, 0.0, -9.81])
    T1 = np.vstack([orc.imu_get(kf, "T_f_w").reshape(3, 4), [0, 0, 0, 1]])
    R1 = T1[:3, :3]
    dT = np.eye(4)
    dT[:3, :3] = orc.imu_get(cur, "dR").reshape(3, 3)
    dT[:3, 3] = orc.imu_get(cur, "dp") + R1 @ orc.imu_get(kf, "v") * dt + 0.5 * R1 @ g * dt * dt
    out = cur.copy()
    out[15:27] = (np.linalg.inv(dT) @ T1)[:3].reshape(12)
    return out


def _aceinna_run(segments):
    """The loop of predictionWithRotation / predictionWithRotation2 (imu_test.cpp:573-702): constant samples at 200 Hz, the
    pose of every frame re-estimated from the pre-integrated deltas, a new keyframe at the start of every segment."""
    T_i_f = np.eye(4)
    T_i_f[:3, :3] = np.diag([1.0, -1.0, -1.0])                                # :576-577
    T_f_i = np.linalg.inv(T_i_f)
    dt, rate = 0.005, 200.0
    out = []
    kf = last = None
    i_kf = 0
    i = 0
    for gyr, acc, n in segments:
        gyr, acc = np.asarray(gyr, float), np.asarray(acc, float)
        if last is None:
            last = orc.imu_state(acc, gyr, T_f_w=T_f_i[:3].reshape(12), is_kf=True)
        else:
            last[27] = 1.0                                                    # cur_frame->setKeyFrame()
        kf, i_kf = last, i
        for _ in range(n):
            i += 1
            cur = orc.process_imu(last, np.zeros(3), np.zeros(3), dt, eta(rate), rate, acc, gyr)
            last = _estimate_transform(kf, cur, (i - i_kf) * dt)
        T_w_f = np.linalg.inv(np.vstack([orc.imu_get(last, "T_f_w").reshape(3, 4), [0, 0, 0, 1]]))
        out.append((T_i_f[:3, :3].T @ T_w_f[:3, 3], T_i_f[:3, :3].T @ orc.imu_get(last, "v")))
    return out


def test_predictionWithRotation():  # imu_test.cpp:573-656, expected values from the Aceinna gnss-ins-sim trajectories
    (p1, v1), (p2, v2) = _aceinna_run([((0.5, 0, 0), (-1, 0, -9.81), 201), ((0.5, 0.2, 0.04), (-1, 0.05, -9.81), 199)])
    assert np.linalg.norm(p1 - [-0.505, 0.813, 0.1]) < 1e-2                   # :614-618
    assert np.linalg.norm(v1 - [-1, 2.41, 0.4]) < 1e-2                        # :619
    assert np.linalg.norm(p2 - [-2.31, 6.18, 1.62]) < 1e-2                    # :646-650
    assert np.linalg.norm(v2 - [-2.95, 8.91, 3.24]) < 1e-2                    # :651


def test_predictionWithRotation2():  # imu_test.cpp:654-703
    ((p, v),) = _aceinna_run([((0.5, 0.2, 0.04), (-1, 0.05, -9.81), 200)])
    assert np.linalg.norm(p - [-0.82143062, 0.80412303, 0.15357111]) < 1e-2   # :695-699
    assert np.linalg.norm(v - [-1.97810799, 2.38035184, 0.5780088]) < 1e-2    # :700-702


# ----------------------------------------------------------------------------------------------------------------
# residual_test.cpp gradient checks (K = diag(100,100), c = (400,400), identity extrinsics)
# ----------------------------------------------------------------------------------------------------------------
K_TEST = np.array([100.0, 100.0, 400.0, 400.0])  # residual_test.cpp:26-30
I34 = np.eye(3, 4).reshape(12)


def sample_visible(rng, T_f_w):
    while True:
        p = rng.uniform(-1, 1, 3)
        pc = T_f_w[:3, :3] @ p + T_f_w[:3, 3]
        if pc[2] < 0.1:
            continue
        uv = np.array([K_TEST[0] * pc[0] / pc[2] + K_TEST[2], K_TEST[1] * pc[1] / pc[2] + K_TEST[3]])
        if (uv >= 0).all() and (uv <= 800).all():
            return p, uv


@pytest.mark.parametrize("seed", range(5))
def test_PriorResidual(seed):  # residual_test.cpp:66-96
    rng = np.random.default_rng(seed)
    T_prior = T34(rand_rot(rng), rng.uniform(-1, 1, 3))
    sq = 100 * np.ones(6)
    for x0 in (np.zeros(6), rng.normal(0, 0.05, 6)):
        _, J = orc.pose_prior_eval(I34, T_prior, sq, x0)
        Jn = num_jac(lambda x: orc.pose_prior_eval(I34, T_prior, sq, x, jac=False)[0], x0)
        assert abs((J - Jn).sum()) < 1e-5           # the reference's criterion
        assert np.abs(J - Jn).max() < 2e-5          # entry-wise


@pytest.mark.parametrize("seed", range(5))
def test_reprojTest(seed):  # residual_test.cpp:98-126
    rng = np.random.default_rng(100 + seed)
    T = np.eye(4)
    T[:3, :3] = rand_rot(rng)
    T[:3, 3] = rng.uniform(-1, 1, 3)
    p, uv = sample_visible(rng, T)
    Tfw = T[:3].reshape(12)
    for x0 in (np.zeros(9), rng.normal(0, 0.01, 9)):
        r, J6, J3 = orc.reproj_eval(uv, K_TEST, I34, Tfw, p, 1.0, x0[:6], x0[6:])
        f = lambda x: orc.reproj_eval(uv, K_TEST, I34, Tfw, p, 1.0, x[:6], x[6:], jac=False)[0]
        Jn = num_jac(f, x0, 1e-7)
        J = np.hstack([J6, J3])
        assert abs((J - Jn).sum()) < 1e-5 * max(1.0, np.abs(J).max())
        assert np.abs(J - Jn).max() < 1e-6 * max(1.0, np.abs(J).max())
    # residual at the exact projection is zero
    r, _, _ = orc.reproj_eval(uv, K_TEST, I34, Tfw, p, 1.0)
    assert np.abs(r).max() < 1e-9


def test_reproj_failed_projection_quirk():
    """Failed projection zeroes the residual but keeps the Jacobian (BundleAdjustmentCERESAnalytic.h:63-68),
    'outside the image' is tested against 2*cx, 2*cy (Camera.cpp:131-132), 'behind' against z < 0.1 (:128)."""
    p_behind = np.array([0.0, 0.0, 0.05])
    r, J6, J3 = orc.reproj_eval([400, 400], K_TEST, I34, I34, p_behind)
    assert np.array_equal(r, np.zeros(2)) and np.abs(J3).max() > 0
    p_out = np.array([5.0, 0.0, 1.0])  # u = 900 > 2*cx
    r, J6, J3 = orc.reproj_eval([400, 400], K_TEST, I34, I34, p_out)
    assert np.array_equal(r, np.zeros(2)) and np.abs(J6).max() > 0
    p_in = np.array([3.9, 0.0, 1.0])   # u = 790 <= 800
    r, _, _ = orc.reproj_eval([400, 400], K_TEST, I34, I34, p_in)
    assert r[0] == pytest.approx(390.0)


@pytest.mark.parametrize("seed", range(5))
def test_angular_factor_gradient(seed):
    """AngularErrCeres_pointxd_dx has NO reference test; same gradient criterion applied to it."""
    rng = np.random.default_rng(200 + seed)
    T = np.eye(4)
    T[:3, :3] = rand_rot(rng)
    T[:3, 3] = rng.uniform(-1, 1, 3)
    Ts = np.eye(4)
    Ts[:3, :3] = rand_rot(rng)
    Ts[:3, 3] = rng.uniform(-0.1, 0.1, 3)
    p = rng.uniform(-3, 3, 3)
    ts = (Ts @ T @ np.append(p, 1))[:3]
    b = ts / np.linalg.norm(ts) + rng.normal(0, 1e-3, 3)
    b /= np.linalg.norm(b)
    sigma = 1.5 / 458.0
    for x0 in (np.zeros(9), rng.normal(0, 0.01, 9)):
        r, J6, J3 = orc.angular_eval(b, Ts[:3].reshape(12), T[:3].reshape(12), p, sigma, x0[:6], x0[6:])
        f = lambda x: orc.angular_eval(b, Ts[:3].reshape(12), T[:3].reshape(12), p, sigma, x[:6], x[6:], jac=False)[0]
        Jn = num_jac(f, x0, 1e-6)
        J = np.hstack([J6, J3])
        assert np.abs(J - Jn).max() < 1e-6 * np.abs(J).max()
    # tangent-basis switch when the bearing is (1,0,0) (AngularAdjustmentCERESAnalytic.h:67-74)
    r, _, _ = orc.angular_eval([1.0, 0, 0], I34, I34, [2.0, 0.0, 0.0], 1.0)
    assert np.abs(r).max() < 1e-15
    r, _, _ = orc.angular_eval([1.0, 0, 0], I34, I34, [2.0, 0.2, 0.0], 1.0)
    assert np.isfinite(r).all() and np.abs(r).max() > 0


@pytest.mark.parametrize("seed", range(3))
def test_PoseToLandmarkResidual(seed):  # residual_test.cpp:236-274
    rng = np.random.default_rng(300 + seed)
    R, t = rand_rot(rng), rng.uniform(-1, 1, 3)
    t_w_lmk = rng.uniform(-1, 1, 3)
    delta = R @ t_w_lmk + t
    sq = (100 * np.eye(3)).reshape(9)
    lmk = t_w_lmk + 0.01 * rng.uniform(-1, 1, 3)
    for x0 in (np.zeros(9), rng.normal(0, 0.05, 9)):
        _, J6, J3 = orc.p2l_eval(delta, T34(R, t), lmk, sq, x0[:6], x0[6:])
        f = lambda x: orc.p2l_eval(delta, T34(R, t), lmk, sq, x[:6], x[6:], jac=False)[0]
        Jn = num_jac(f, x0, 1e-6)
        J = np.hstack([J6, J3])
        assert abs((J - Jn).sum()) < 1e-5
        assert np.abs(J - Jn).max() < 1e-5


def test_simuEuroc_integration_on_the_committed_slice():  # imu_test.cpp:760-806
    """processIMU dead-reckoning against the test's own "simple integration", every step within 0.1 in Frobenius norm; run on the
    committed slice of euroc_gt.csv (samples 2000..4929) instead of the whole 28 k-sample file."""
    from tests import ref_fixtures as rf

    acc, gyr, R, p, v, ts_f, ts_ns = rf._euroc_samples()
    g = np.array([0.0, 0.0, -9.81])
    T0 = rf._T_f_w(R[0], p[0])
    kf = orc.imu_state(acc[0], gyr[0], T_f_w=T0, v=v[0], is_kf=True)
    last = kf
    Rp, tp, vp = R[0].copy(), p[0].copy(), v[0].copy()
    worst = 0.0
    for i in range(1, acc.shape[0]):
        dt = (ts_f[i] - ts_f[i - 1]) * 1e-9
        last = orc.process_imu(last, np.zeros(3), np.zeros(3), (int(ts_ns[i]) - int(ts_ns[i - 1])) * 1e-9, eta(200.0), 200.0, acc[i], gyr[i])
        vn = vp + g * dt + Rp @ acc[i - 1] * dt                                  # :789-795
        tp = tp + vp * dt + 0.5 * g * dt * dt + 0.5 * Rp @ acc[i - 1] * dt * dt
        Rp = Rp @ orc.exp_so3(gyr[i - 1] * dt)
        vp = vn
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = Rp, tp
        T_f_w = np.vstack([orc.imu_get(last, "T_f_w").reshape(3, 4), [0, 0, 0, 1]])
        worst = max(worst, np.linalg.norm(T @ T_f_w - np.eye(4)))
    assert worst < 0.1                                                           # :799-800
    assert worst < 1e-6      # the two integrations are the same recursion; only the timestamp rounding differs

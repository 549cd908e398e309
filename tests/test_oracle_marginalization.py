"""Oracle groundwork for SURVEY.md §8 row a15 (marginal-prior construction, "next"): the dense core of the reference's
Marginalization (computeInformationAndGradient / computeSchurComplement / rankReveallingDecomposition /
computeJacobiansAndResiduals, cpp/src/optimizers/marginalization.cpp:145-265,318-342,516-530) restated in oracle/marg.hpp and
pinned on the reference's own toy graph (cpp/tests/marginalization_test.cpp:14-24 setup, :219-223 sizes, :300-313 checks)."""
import numpy as np
import pytest

from oracle import oracle

K = np.array([100.0, 100.0, 400.0, 400.0])                       # marginalization_test.cpp:35-39
T_ID = np.hstack([np.eye(3), np.zeros((3, 1))]).reshape(-1)        # frame 0 at the origin, left camera on the frame
T_RIGHT = np.hstack([np.eye(3), np.array([[0.0], [0.2], [0.0]])]).reshape(-1)  # right camera: translation y = 0.2 (:49-51)
LMKS = [np.array([0.5, 0.0, 2.0]), np.array([-1.0, 0.0, 2.0]), np.array([1.0, 0.0, 2.0])]  # :79, :90, :100
# The toy graph is singular by construction: with two kept landmarks the marginalised rig (and its lonely landmark) can still
# rotate about the line through them, so A_mm has one eigenvalue that is zero up to rounding (~1e-11) — noise around the
# reference's 1e-12 threshold.  The reference test only asserts sizes / symmetry / a non-zero off-diagonal block for it; the
# numerical invariants are checked on the same graph with a third kept landmark (no gauge freedom left).
LMKS_GENERIC = [np.array([0.5, 0.3, 2.0]), np.array([-1.0, -0.2, 2.6]), np.array([1.0, 0.25, 1.7]), np.array([0.1, -0.6, 3.1])]


def project(T_s_f, p):
    T = T_s_f.reshape(3, 4)
    pc = T[:, :3] @ p + T[:, 3]
    return np.array([K[0] * pc[0] / pc[2] + K[2], K[1] * pc[1] / pc[2] + K[3]])


def information(noise=0.0, seed=0, lmks=LMKS):
    """A, b of the six stereo reprojection factors attached to frame 0 (computeInformationAndGradient, :145-211) with the index
    map the reference test asserts: frame 0 -> 0, l0 (marginalised) -> 6, l1 -> 9, l2 -> 12 (marginalization_test.cpp:300-303)."""
    rng = np.random.default_rng(seed)
    idx_lmk = {l: 6 + 3 * l for l in range(len(lmks))}
    N = 6 + 3 * len(lmks)
    A, b = np.zeros((N, N)), np.zeros(N)
    for l, p in enumerate(lmks):
        for T_s_f in (T_ID, T_RIGHT):
            uv = project(T_s_f, p) + noise * rng.normal(size=2)
            r, J6, J3 = oracle.reproj_eval(uv, K, T_s_f, T_ID, p)  # ReprojectionErrCeres_pointxd_dx, sigma = 1
            blocks = ((0, J6), (idx_lmk[l], J3))
            for i, (ii, Ji) in enumerate(blocks):
                for jj, Jj in blocks[i:]:
                    A[ii:ii + Ji.shape[1], jj:jj + Jj.shape[1]] += Ji.T @ Jj
                    if jj != ii:
                        A[jj:jj + Jj.shape[1], ii:ii + Ji.shape[1]] = A[ii:ii + Ji.shape[1], jj:jj + Jj.shape[1]].T
                b[ii:ii + Ji.shape[1]] += Ji.T @ r
    return A, b


def numpy_reference(A, b, m, eps=1e-12):
    """The same algebra with numpy.linalg.eigh (independent of the oracle's Jacobi eigen-solver)."""
    Amm = 0.5 * (A[:m, :m] + A[:m, :m].T)
    w, V = np.linalg.eigh(Amm)
    Ainv = (V * np.where(w > eps, 1.0 / np.where(w > eps, w, 1.0), 0.0)) @ V.T
    Arm, Arr = A[m:, :m], A[m:, m:]
    return Arr - Arm @ Ainv @ Arm.T, b[m:] - Arm @ Ainv @ b[:m]


def test_reference_toy_graph_sizes_and_symmetry():
    A, b = information()
    out = oracle.schur_prior(A, b, m=9)               # _m = 9 (pose 6 + lonely landmark 3), _n = 6 (:219-220)
    assert out is not None
    Ak = out["Ak"]
    assert Ak.size == 36                               # ASSERT_EQ(_marg._Ak.size(), 36)                     (:309)
    assert np.linalg.norm(Ak - Ak.T) < 1e-8            # ASSERT_NEAR((_Ak - _Ak^T).norm(), 0, 1e-8)          (:310)
    assert abs(np.trace(Ak[0:3, 3:6])) > 0             # computeOffDiag(l1, l2) > 0, l1 -> 0, l2 -> 3        (:305-306, :313)
    assert np.allclose(b, 0) and np.allclose(out["bk"], 0)  # perfect projections: zero gradient


def test_fewer_than_four_kept_parameters_fails_like_the_reference():
    A, b = information()
    assert oracle.schur_prior(A[:12, :12], b[:12], m=9) is None   # n = 3 < 4 -> computeSchurComplement() == false (:215, test :334)


@pytest.mark.parametrize("noise", [0.0, 0.7])
def test_schur_complement_and_prior_factor_invariants(noise):
    A, b = information(noise=noise, seed=3, lmks=LMKS_GENERIC)
    assert np.linalg.eigvalsh(A[:9, :9])[0] > 1e-3
    out = oracle.schur_prior(A, b, m=9)
    Ak_np, bk_np = numpy_reference(A, b, 9)
    scale = np.abs(Ak_np).max()
    assert np.abs(out["Ak"] - Ak_np).max() <= 1e-9 * scale
    assert np.abs(out["bk"] - bk_np).max() <= 1e-9 * max(1.0, np.abs(bk_np).max())
    # rank-revealing decomposition: eigenvalues above 1e-12 kept, ascending, orthonormal U.  Once frame 0 is gone the kept
    # landmarks have a 6-dof gauge freedom: Ak has rank 3, its other eigenvalues are rounding noise (|.| ~ 1e-10) on either
    # side of the reference's absolute threshold — how many of THOSE are kept is solver noise, so only the significant part
    # of the spectrum is compared.
    lam = np.linalg.eigvalsh(Ak_np)
    big = lam[lam > 1e-6 * scale]
    assert big.size == 3 and out["n_full"] >= 3
    assert np.all(out["Lambda"] > 1e-12) and np.all(np.diff(out["Lambda"]) >= 0)
    assert np.allclose(out["Lambda"][-3:], big, rtol=1e-9)
    assert np.all(out["Lambda"][:-3] < 1e-8 * scale)
    U = out["U"]
    assert np.abs(U.T @ U - np.eye(out["n_full"])).max() < 1e-10
    # the prior factor r = r0 + J dx reproduces the marginal information and gradient: J^T J = Ak (on its range), J^T r0 = -bk
    J, r0 = out["J"], out["r0"]
    assert J.shape == (out["n_full"], 9)
    assert np.abs(J.T @ J - Ak_np).max() <= 1e-8 * scale
    assert np.abs(J.T @ r0 + bk_np).max() <= 1e-8 * max(1.0, np.abs(bk_np).max())


def test_rank_deficient_marginal_block_uses_the_pseudo_inverse():
    """A marginalised parameter without information (zero row/column) must be ignored by the eps-thresholded inverse (:234-240)."""
    A, b = information(noise=0.3, seed=5, lmks=LMKS_GENERIC)
    A2, b2 = np.zeros((19, 19)), np.zeros(19)
    A2[1:, 1:], b2[1:] = A, b                        # one more marginalised parameter in front, all zeros
    o1, o2 = oracle.schur_prior(A, b, m=9), oracle.schur_prior(A2, b2, m=10)
    assert np.abs(o1["Ak"] - o2["Ak"]).max() <= 1e-9 * np.abs(o1["Ak"]).max()
    assert np.abs(o1["bk"] - o2["bk"]).max() <= 1e-9 * max(1.0, np.abs(o1["bk"]).max())


@pytest.mark.parametrize("n,rank", [(1, 1), (2, 2), (7, 7), (40, 40), (60, 23), (150, 150)])
def test_symmetric_eigen_solvers(n, rank):
    """The two eigen-solvers of marg.hpp (Householder + implicit QL — the one schur_prior uses, the same two stages as Eigen's
    SelfAdjointEigenSolver — and cyclic Jacobi) against numpy.linalg.eigvalsh, on full-rank and rank-deficient matrices."""
    rng = np.random.default_rng(n * 100 + rank)
    B = rng.normal(size=(n, rank))
    A = B @ B.T + (0.0 if rank < n else 1e-3) * np.eye(n)
    ref = np.linalg.eigvalsh(A)
    scale = max(1.0, np.abs(ref).max())
    for method in ("ql", "jacobi"):
        w, V = oracle.sym_eig(A, method)
        assert np.all(np.diff(w) >= 0)                                   # ascending, as Eigen returns them
        assert np.abs(w - ref).max() <= 1e-12 * scale * n
        assert np.abs(V.T @ V - np.eye(n)).max() <= 1e-12 * n
        assert np.abs(V @ np.diag(w) @ V.T - A).max() <= 1e-12 * scale * n

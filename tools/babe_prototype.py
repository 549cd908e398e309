"""Numpy prototype of the "burn at both ends" (two-way dissection) variant of the banded block Cholesky solve planned for
k_chol_band: the block rows are split into a left half eliminated top-down, a right half eliminated bottom-up
(the same kernel body on the index-reversed matrix) and a separator of `bw` blocks in between that receives the trailing
updates of both halves.  There is no extra fill — the sequential chain of block steps drops from nb to ceil((nb - bw) / 2) + bw
— and both triangular solves split the same way.

This file only checks the algebra and the index mapping (reversal at 16-column block granularity, what the second CTA must
zero-initialise, what crosses between the CTAs) against numpy.linalg.solve on random banded SPD systems of the shapes the
window solve produces (C3: nb = 46, bw = 3; C5: nb = 188, bw = 3..4).  It is not on any product path.

    python tools/babe_prototype.py
"""
import numpy as np

BN = 16


def banded_spd(nb, bw, rng):
    n = nb * BN
    A = np.zeros((n, n))
    for i in range(nb):
        for j in range(max(0, i - bw), i + 1):
            B = rng.normal(size=(BN, BN))
            A[i * BN:(i + 1) * BN, j * BN:(j + 1) * BN] = B
            A[j * BN:(j + 1) * BN, i * BN:(i + 1) * BN] = B.T
    A = A @ A.T                                   # bandwidth 2 bw in blocks ...
    for i in range(nb):                           # ... cut back to bw and made diagonally dominant
        for j in range(nb):
            if abs(i - j) > bw:
                A[i * BN:(i + 1) * BN, j * BN:(j + 1) * BN] = 0
    A += np.eye(n) * (np.abs(A).sum(axis=1).max() + 1.0)
    return A


def blk(M, i, j):
    return M[i * BN:(i + 1) * BN, j * BN:(j + 1) * BN]


def eliminate(W, g, steps, nb, bw):
    """Right-looking banded block Cholesky of the first `steps` block columns of the working copy W (lower part used), forward
    substitution of g alongside: exactly what the chain / row-solve / update warps of k_chol_band do per block step."""
    for k in range(steps):
        Lkk = np.linalg.cholesky(blk(W, k, k))
        blk(W, k, k)[:] = Lkk
        g[k * BN:(k + 1) * BN] = np.linalg.solve(Lkk, g[k * BN:(k + 1) * BN])
        hi = min(nb, k + bw + 1)
        for i in range(k + 1, hi):
            blk(W, i, k)[:] = np.linalg.solve(Lkk, blk(W, i, k).T).T          # L_ik = A_ik L_kk^-T
            g[i * BN:(i + 1) * BN] -= blk(W, i, k) @ g[k * BN:(k + 1) * BN]
        for i in range(k + 1, hi):
            for j in range(k + 1, i + 1):
                blk(W, i, j)[:] -= blk(W, i, k) @ blk(W, j, k).T
    return W, g


def back_substitute(W, y, x, first, count, nb, bw):
    """x_k = L_kk^-T (y_k - sum_d L_(k+d,k)^T x_(k+d)) for k = first + count - 1 .. first."""
    for k in range(first + count - 1, first - 1, -1):
        r = y[k * BN:(k + 1) * BN].copy()
        for i in range(k + 1, min(nb, k + bw + 1)):
            r -= blk(W, i, k).T @ x[i * BN:(i + 1) * BN]
        x[k * BN:(k + 1) * BN] = np.linalg.solve(blk(W, k, k).T, r)
    return x


def solve_babe(A, b, nb, bw):
    n = nb * BN
    nl = (nb - bw + 1) // 2                      # left interior: blocks 0 .. nl-1            (CTA 0, top-down)
    nr = nb - bw - nl                            # right interior: blocks nb-nr .. nb-1        (CTA 1, bottom-up)
    rev = np.arange(n)[::-1]                     # CTA 1 works on P A P, P = full index reversal: block q <-> nb-1-q,
    # and inside a block column c <-> 15-c, so 16-column blocks stay aligned because n is a multiple of 16
    # ---- CTA 0: its own copy of rows 0 .. nl+bw-1 (interior + separator), separator entries as in A
    W0, g0 = A.copy(), b.copy()
    eliminate(W0, g0, nl, nb, bw)
    # ---- CTA 1: reversed matrix, rows 0 .. nr+bw-1 of it; the separator-separator blocks and the separator part of the
    # right-hand side start from ZERO so that what it ships is only its own trailing update
    A1, b1 = A[np.ix_(rev, rev)].copy(), b[rev].copy()
    sep1 = slice(nr * BN, (nr + bw) * BN)
    A1[sep1, sep1] = 0.0
    b1[sep1] = 0.0
    W1, g1 = A1, b1
    eliminate(W1, g1, nr, nb, bw)
    # ---- hand-over (DSMEM): bw (bw + 1) / 2 blocks of the lower triangle + bw * 16 right-hand-side entries, index-reversed
    sep0 = slice(nl * BN, (nl + bw) * BN)
    dS = W1[sep1, sep1]
    dS = np.tril(dS) + np.tril(dS, -1).T         # CTA 1 only maintains its lower triangle
    W0[sep0, sep0] += dS[::-1, ::-1]
    g0[sep0] += g1[sep1][::-1]
    # ---- CTA 0 continues its chain over the separator: bw more block steps (rows beyond the separator are not its business)
    Ws, gs = W0[:(nl + bw) * BN, :(nl + bw) * BN], g0[:(nl + bw) * BN]
    for k in range(nl, nl + bw):
        Lkk = np.linalg.cholesky(blk(Ws, k, k))
        blk(Ws, k, k)[:] = Lkk
        gs[k * BN:(k + 1) * BN] = np.linalg.solve(Lkk, gs[k * BN:(k + 1) * BN])
        for i in range(k + 1, nl + bw):
            blk(Ws, i, k)[:] = np.linalg.solve(Lkk, blk(Ws, i, k).T).T
            gs[i * BN:(i + 1) * BN] -= blk(Ws, i, k) @ gs[k * BN:(k + 1) * BN]
        for i in range(k + 1, nl + bw):
            for j in range(k + 1, i + 1):
                blk(Ws, i, j)[:] -= blk(Ws, i, k) @ blk(Ws, j, k).T
    # ---- backward: separator on CTA 0, its bw * 16 solution entries go to CTA 1, then both interiors in parallel
    x = np.zeros(n)
    back_substitute(Ws, gs, x[:(nl + bw) * BN], nl, bw, nl + bw, bw)
    x1 = np.zeros(n)
    x1[sep1] = x[sep0][::-1]
    back_substitute(W1, g1, x1, 0, nr, nr + bw, bw)
    back_substitute(Ws, gs, x[:(nl + bw) * BN], 0, nl, nl + bw, bw)
    x[(nl + bw) * BN:] = x1[:nr * BN][::-1]
    return x, nl, nr


def main():
    rng = np.random.default_rng(7)
    for nb, bw in ((47, 3), (47, 4), (13, 3), (8, 3), (7, 3), (187, 3), (5, 1), (3, 1)):
        A = banded_spd(nb, bw, rng)
        b = rng.normal(size=nb * BN)
        x, nl, nr = solve_babe(A, b, nb, bw)
        ref = np.linalg.solve(A, b)
        err = np.abs(x - ref).max() / np.abs(ref).max()
        print(f"nb={nb:4d} bw={bw}: left {nl:3d} + right {nr:3d} + separator {bw}: sequential block steps {nb} -> {max(nl, nr) + bw}, "
              f"max rel err {err:.1e}")
        assert err < 1e-10
    print("ok")


if __name__ == "__main__":
    main()

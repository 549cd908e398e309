"""Numpy prototype of the P-way dissection of the banded block Cholesky solve — the next step for k_chol_band at the 200-keyframe
configuration (tools/babe_prototype.py is the two-way model that shipped in round 2).

The nb block rows are cut into P interiors separated by P - 1 separators of `bw` block rows each.  Every interior is eliminated by
its own CTA: the two END parts exactly as today (top-down / bottom-up on the reversed matrix, no extra fill); a MIDDLE part is
eliminated top-down and drags a BORDER along — the coupling of its rows to the separator above it fills in (a dense column of bw
blocks per interior block row: the "bordered band"), and its trailing updates land on three places: the separator above, the
separator below, and a NEW coupling block between the two.  What is left is a block-tridiagonal system over the separators
((P - 1) bw block rows, half-bandwidth 2 bw - 1 blocks) that one CTA finishes; then every part back-substitutes its interior.

This file checks the algebra against numpy.linalg.solve, verifies the structural claims (where fill appears, what crosses between
CTAs) and counts the sequential block steps:  max interior + separator chain, against nb for one CTA and nl + bw for two.

    python tools/pway_prototype.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from babe_prototype import BN, banded_spd  # noqa: E402  (same directory)


def partition(nb, bw, P):
    """Interior block ranges [a, b) and separator block ranges of a P-way split (interiors as equal as possible)."""
    n_int = nb - (P - 1) * bw
    assert n_int >= P, "band too short for this many parts"
    base, extra = divmod(n_int, P)
    interiors, seps, pos = [], [], 0
    for p in range(P):
        m = base + (1 if p < extra else 0)
        interiors.append((pos, pos + m))
        pos += m
        if p < P - 1:
            seps.append((pos, pos + bw))
            pos += bw
    return interiors, seps


def solve_pway(A, b, nb, bw, P):
    n = nb * BN
    interiors, seps = partition(nb, bw, P)
    rows = lambda r: slice(r[0] * BN, r[1] * BN)
    sep_idx = np.concatenate([np.arange(s[0] * BN, s[1] * BN) for s in seps])
    S = A[np.ix_(sep_idx, sep_idx)].copy()          # separator system, updated by every part
    gS = b[sep_idx].copy()
    sep_off = {k: k * bw * BN for k in range(len(seps))}
    factors, fill_blocks, chain = [], 0, []
    for p, I in enumerate(interiors):
        sl = rows(I)
        AII = A[sl, sl]
        L = np.linalg.cholesky(AII)                 # banded: the factor of a band stays inside the band
        # boundary of this part: the separator above (p - 1) and below (p); an end part has one of them
        bnd = [k for k in (p - 1, p) if 0 <= k < len(seps)]
        cols = np.concatenate([np.arange(seps[k][0] * BN, seps[k][1] * BN) for k in bnd])
        AIB = A[sl, :][:, cols]
        Y = np.linalg.solve(L, AIB)                 # L^-1 A_IB: THE BORDER.  Its sparsity is the fill the kernel must hold:
        m = I[1] - I[0]
        for q, k in enumerate(bnd):
            Yk = Y[:, q * bw * BN:(q + 1) * bw * BN]
            nz_rows = [r for r in range(m) if np.abs(Yk[r * BN:(r + 1) * BN]).max() > 1e-13]
            if k == p:                              # separator BELOW: touched by the last bw block rows only (ordinary band)
                assert all(r >= m - bw for r in nz_rows)
            elif p < P - 1:                         # separator ABOVE a top-down elimination: the coupling fills every row below
                fill_blocks += max(0, len(nz_rows) - bw) * bw   # (the LAST part runs bottom-up on the reversed matrix: no fill)
        y = np.linalg.solve(L, b[sl])
        # trailing updates on the separator system: diagonal blocks of the two separators and — middle parts — their coupling
        for qa, ka in enumerate(bnd):
            Ya = Y[:, qa * bw * BN:(qa + 1) * bw * BN]
            oa = sep_off[ka]
            gS[oa:oa + bw * BN] -= Ya.T @ y
            for qb, kb in enumerate(bnd):
                Yb = Y[:, qb * bw * BN:(qb + 1) * bw * BN]
                ob = sep_off[kb]
                S[oa:oa + bw * BN, ob:ob + bw * BN] -= Ya.T @ Yb
        factors.append((sl, L, Y, y, bnd))
        chain.append(m)
    # the separator system is block tridiagonal in separators (coupling only between neighbours: created by the middle parts)
    ns = len(seps)
    for a in range(ns):
        for c in range(ns):
            if abs(a - c) > 1:
                assert np.abs(S[sep_off[a]:sep_off[a] + bw * BN, sep_off[c]:sep_off[c] + bw * BN]).max() < 1e-9
    xS = np.linalg.solve(S, gS)                     # one CTA: a banded Cholesky of (P - 1) bw block rows
    x = np.zeros(n)
    x[sep_idx] = xS
    for sl, L, Y, y, bnd in factors:                # every part back-substitutes its interior in parallel
        xb = np.concatenate([xS[sep_off[k]:sep_off[k] + bw * BN] for k in bnd])
        x[sl] = np.linalg.solve(L.T, y - Y @ xb)
    steps = max(chain) + (P - 1) * bw
    return x, steps, fill_blocks, interiors, seps


def main():
    rng = np.random.default_rng(3)
    for nb, bw in ((188, 4), (188, 3), (47, 3), (47, 4)):
        A = banded_spd(nb, bw, rng)
        b = rng.normal(size=nb * BN)
        ref = np.linalg.solve(A, b)
        line = [f"nb={nb:4d} bw={bw}: 1 CTA {nb} steps, 2 CTAs {(nb - bw + 1) // 2 + bw}"]
        for P in (3, 4, 6, 8):
            if nb - (P - 1) * bw < P:
                continue
            x, steps, fill, interiors, seps = solve_pway(A, b, nb, bw, P)
            err = np.abs(x - ref).max() / np.abs(ref).max()
            assert err < 1e-9, err
            line.append(f"P={P}: {steps} steps (+{fill} fill blocks)")
        print(", ".join(line))
    print("ok")


if __name__ == "__main__":
    main()

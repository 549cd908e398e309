"""Element-level emulation of the index arithmetic of the -DSDV_BAND_BABE=1 path of k_chol_band (sdv_chol_band.cuh): what
each CTA loads (stored lower triangle, reversed reads, zero-filled separator), the local order of damping / right-hand side,
the anti-transposed hand-over of CTA 1's separator update through the window slots, the exchange of x and the scatter into
dxp.  The block arithmetic itself (chain, row solves, updates) is replaced by numpy; only the addressing is the kernel's.

    python tools/babe_index_emulation.py
"""
import numpy as np

BN = 16


def run(nbg, bw, n, seed=0):
    rng = np.random.default_rng(seed)
    n_pad = nbg * BN
    assert 0 <= n_pad - n < 2 * BN      # the library pads to a multiple of 32
    # reduced system as the kernel sees it: lower triangle stored (upper part = NaN: must never be used), padding rows zero
    S = np.zeros((n_pad, n_pad))
    for i in range(nbg):
        for j in range(max(0, i - bw), i + 1):
            B = rng.normal(size=(BN, BN)) * 0.1
            S[i * BN:(i + 1) * BN, j * BN:(j + 1) * BN] = B
    S = np.tril(S)
    S = S + np.tril(S, -1).T
    S[n:, :] = 0
    S[:, n:] = 0
    damp = np.where(np.arange(n_pad) < n, 3.0 + rng.random(n_pad), -1.0)      # dmp: LM damping, negative = padding column
    g = np.where(np.arange(n_pad) < n, rng.normal(size=n_pad), 0.0)
    A = np.where(np.tril(np.ones((n_pad, n_pad))) > 0, S, np.nan)            # what global memory holds
    full = S + np.diag(np.where(damp < 0, 1.0, damp))                        # the system the kernel solves: (S + D) x = g
    xref = np.linalg.solve(full, g)

    nl = (nbg - bw + 1) // 2
    nr = nbg - bw - nl
    R, bwp = bw + 3, bw + 1
    cta = []
    for rev in (False, True):
        nint = nr if rev else nl
        nb = nint + bw
        W = np.zeros((nb * BN, nb * BN))           # local system, lower blocks (diagonal blocks full), local order
        gs, dmp = np.zeros(nb * BN), np.zeros(nb * BN)
        # band_sysprep: local order of damping and right-hand side
        for i in range(n_pad):
            il = n_pad - 1 - i if rev else i
            if il >= nb * BN:
                continue
            sep = rev and il >= nint * BN
            dmp[il] = 0.0 if sep else damp[i]
            gs[il] = 0.0 if sep else g[i]
        # load_row + fix_row
        for i in range(nb):
            for j in range(max(0, i - bw), i + 1):
                for r in range(BN):
                    for c in range(BN):
                        if rev:
                            if i >= nint and j >= nint:
                                v = 0.0
                            else:
                                gr, gc = n_pad - 1 - (i * BN + r), n_pad - 1 - (j * BN + c)
                                v = A[max(gr, gc), min(gr, gc)]
                        else:
                            v = A[i * BN + r, j * BN + c]  # may be NaN in the upper part of a diagonal block
                        W[i * BN + r, j * BN + c] = v
            for r in range(BN):
                d = dmp[i * BN + r]
                W[i * BN + r, i * BN + r] = 1.0 if d < 0 else W[i * BN + r, i * BN + r] + d
        cta.append(dict(rev=rev, nint=nint, nb=nb, W=W, gs=gs))

    def blk(W, i, j):
        return W[i * BN:(i + 1) * BN, j * BN:(j + 1) * BN]

    def lower_sym(B):   # the chain only uses the lower triangle of a diagonal block
        L = np.tril(B)
        return L + np.tril(L, -1).T

    def steps(c, k0, k1):
        W, gs, nb = c["W"], c["gs"], c["nb"]
        for k in range(k0, k1):
            Lkk = np.linalg.cholesky(lower_sym(blk(W, k, k)))
            blk(W, k, k)[:] = Lkk
            gs[k * BN:(k + 1) * BN] = np.linalg.solve(Lkk, gs[k * BN:(k + 1) * BN])
            hi = min(nb, k + bw + 1)
            for i in range(k + 1, hi):
                blk(W, i, k)[:] = np.linalg.solve(Lkk, blk(W, i, k).T).T
                gs[i * BN:(i + 1) * BN] -= blk(W, i, k) @ gs[k * BN:(k + 1) * BN]
            for i in range(k + 1, hi):
                for j in range(k + 1, i + 1):
                    if i == j:
                        blk(W, i, i)[:] = np.where(np.isnan(blk(W, i, i)), 0.0, blk(W, i, i))  # never-used upper garbage
                    blk(W, i, j)[:] -= blk(W, i, k) @ blk(W, j, k).T

    c0, c1 = cta
    steps(c0, 0, c0["nint"])
    steps(c1, 0, c1["nint"])
    # hand-over, exactly the kernel's formula (window slots collapse to the block index here)
    nint = c0["nint"]
    for a in range(bw):
        for b in range(a + 1):
            i, j = nint + a, nint + b
            i1, j1 = nbg - 1 - j, nbg - 1 - i
            for r in range(BN):
                for c in range(BN):
                    src = blk(c1["W"], i1, j1)[15 - c, 15 - r]
                    if np.isnan(blk(c0["W"], i, j)[r, c]):
                        continue      # upper part of CTA 0's own diagonal block: garbage in, never used
                    blk(c0["W"], i, j)[r, c] += src
    for t in range(bw * BN):
        c0["gs"][nint * BN + t] += c1["gs"][n_pad - 1 - (nint * BN + t)]
    steps(c0, nint, c0["nb"])

    # backward solves; x exchange with the kernel's addresses
    dxp = np.zeros(n_pad)

    def backward(c, kstart, other):
        W, gs, nb, rev = c["W"], c["gs"], c["nb"], c["rev"]
        for k in range(kstart, -1, -1):
            rv = gs[k * BN:(k + 1) * BN].copy()
            for d in range(1, min(bw, nb - 1 - k) + 1):
                rv -= blk(W, k + d, k).T @ gs[(k + d) * BN:(k + d + 1) * BN]
            x = np.linalg.solve(np.tril(blk(W, k, k)).T, rv)
            gs[k * BN:(k + 1) * BN] = x
            for cc in range(BN):
                p = k * BN + cc
                dxp[n_pad - 1 - p if rev else p] = -x[cc]
                if rev or k >= c["nint"]:
                    other["gs_remote"][n_pad - 1 - p] = x[cc]

    c0["gs_remote"] = np.zeros(n_pad)      # CTA 0's gs has n_pad entries (the epilogue reads all of them)
    c0["gs_remote"][:c0["nb"] * BN] = 0
    c1["gs_remote"] = c1["gs"]             # stores into CTA 1's gs land in its local array
    big0 = np.zeros(n_pad)
    c0_store = dict(gs_remote=big0)        # CTA 1 -> CTA 0
    # separator first on CTA 0 (k = nb-1 .. nint), its x goes to CTA 1; then both interiors
    W0, gs0 = c0["W"], c0["gs"]
    for k in range(c0["nb"] - 1, nint - 1, -1):
        rv = gs0[k * BN:(k + 1) * BN].copy()
        for d in range(1, min(bw, c0["nb"] - 1 - k) + 1):
            rv -= blk(W0, k + d, k).T @ gs0[(k + d) * BN:(k + d + 1) * BN]
        x = np.linalg.solve(np.tril(blk(W0, k, k)).T, rv)
        gs0[k * BN:(k + 1) * BN] = x
        for cc in range(BN):
            p = k * BN + cc
            dxp[p] = -x[cc]
            c1["gs"][n_pad - 1 - p] = x[cc]                      # dsmem_store(gs + (n_pad-1-p), rank 1, x)
    backward(c1, c1["nint"] - 1, c0_store)
    for k in range(nint - 1, -1, -1):
        rv = gs0[k * BN:(k + 1) * BN].copy()
        for d in range(1, min(bw, c0["nb"] - 1 - k) + 1):
            rv -= blk(W0, k + d, k).T @ gs0[(k + d) * BN:(k + d + 1) * BN]
        x = np.linalg.solve(np.tril(blk(W0, k, k)).T, rv)
        gs0[k * BN:(k + 1) * BN] = x
        dxp[k * BN:(k + 1) * BN] = -x
    # epilogue view of CTA 0: its own rows + what CTA 1 stored remotely
    x_epi = big0.copy()
    x_epi[:c0["nb"] * BN] = gs0
    err_dxp = np.abs(-dxp - xref).max() / np.abs(xref).max()
    err_epi = np.abs(x_epi - xref).max() / np.abs(xref).max()
    return err_dxp, err_epi, nl, nr


def main():
    for nbg, bw, n in ((20, 3, 20 * 16 - 5), (21, 3, 21 * 16), (24, 4, 24 * 16 - 15), (46, 3, 735), (48, 3, 748)):
        e1, e2, nl, nr = run(nbg, bw, n)
        print(f"nb={nbg} bw={bw} n={n}: nl={nl} nr={nr}  dxp error {e1:.1e}  epilogue-x error {e2:.1e}")
        assert e1 < 1e-10 and e2 < 1e-10
    print("ok")


if __name__ == "__main__":
    main()

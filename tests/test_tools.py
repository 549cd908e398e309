"""The planning prototypes under tools/ stay runnable (they are cited by DESIGN.md §7)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_two_way_dissection_prototype_matches_a_direct_solve():
    import numpy as np

    babe = _load("babe_prototype")
    rng = np.random.default_rng(11)
    for nb, bw in ((47, 3), (9, 2), (4, 1)):
        A = babe.banded_spd(nb, bw, rng)
        b = rng.normal(size=nb * babe.BN)
        x, nl, nr = babe.solve_babe(A, b, nb, bw)
        assert nl + nr + bw == nb and abs(nl - nr) <= 1
        ref = np.linalg.solve(A, b)
        assert np.abs(x - ref).max() <= 1e-10 * np.abs(ref).max()


def test_two_way_dissection_index_emulation():
    """Element-level emulation of the addressing of the -DSDV_BAND_BABE=1 kernel path (loads from the stored triangle, reversed
    local orders, anti-transposed hand-over, exchange of x, scatter into dxp) against a direct solve."""
    emu = _load("babe_index_emulation")
    for nbg, bw, n in ((20, 3, 20 * 16 - 5), (24, 4, 24 * 16 - 15)):
        e_dxp, e_epi, nl, nr = emu.run(nbg, bw, n)
        assert nl + nr + bw == nbg and e_dxp < 1e-10 and e_epi < 1e-10


def test_p_way_dissection_prototype_matches_a_direct_solve():
    """tools/pway_prototype.py: P interiors + P - 1 separators, bordered middle parts, block-tridiagonal
    separator system — exact against a direct solve, and shorter chains than the two-way split that ships."""
    import numpy as np

    pw = _load("pway_prototype")
    rng = np.random.default_rng(5)
    for nb, bw, P in ((60, 3, 4), (47, 3, 3), (40, 2, 5)):
        A = pw.banded_spd(nb, bw, rng)
        b = rng.normal(size=nb * pw.BN)
        x, steps, fill, interiors, seps = pw.solve_pway(A, b, nb, bw, P)
        assert len(interiors) == P and len(seps) == P - 1 and sum(b_ - a_ for a_, b_ in interiors) + (P - 1) * bw == nb
        ref = np.linalg.solve(A, b)
        assert np.abs(x - ref).max() <= 1e-9 * np.abs(ref).max()
        assert steps < (nb - bw + 1) // 2 + bw and fill > 0

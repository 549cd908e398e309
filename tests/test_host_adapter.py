"""The C++ adapter (sadvio_b200/host/b200_optimizer.hpp) must flatten the pointer graph in the reference's walk order
(addResidualsLocalMap, AngularAdjustmentCERESAnalytic.cpp:212-339; addIMUResiduals, AOptimizer.cpp:22-96) so that the
(landmark, frame, camera) visibility triplets are BIT-EXACT, and write the solution back as AOptimizer.cpp:391-434 does."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from sadvio_b200 import build, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def adapter_exe():
    lib = build.build()
    exe = os.path.join(tempfile.gettempdir(), "sdv_adapter_check")
    src = os.path.join(ROOT, "sadvio_b200", "host", "adapter_check.cpp")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-L", os.path.dirname(lib),
                    "-lsadvio_b200", f"-Wl,-rpath,{os.path.dirname(lib)}"], check=True)
    return exe


def graph_text(win, rng, extras=True):
    """Serialise a synth window as a SaDVIO-style pointer graph (frames oldest -> newest) and compute, independently,
    the flattening the reference walk produces.  `extras` adds everything the reference filters out."""
    F = win.n_frames
    lines = []
    n_extra_frames = 2 if extras else 0       # two frames that are NOT in the window (older, dropped)
    NF = F + n_extra_frames
    lines.append(str(NF))
    order = list(range(F - 1, -1, -1))        # window index (newest first) -> time order
    nonkf = {order[2]} if extras and F > 4 else set()   # a non-keyframe inside the window: its features are skipped (:272)
    frame_time_pos = {}                       # window index f -> position in the file
    pos = 0
    fmt = lambda a: " ".join(repr(float(x)) for x in np.asarray(a).reshape(-1))
    # extra (out-of-window) frames first: they are older
    for e in range(n_extra_frames):
        lines.append(f"{100 + e} 1 0 0 0 2")
        lines.append(fmt(np.eye(3, 4)) + " " + fmt(np.eye(3, 4)) + " " + fmt(np.zeros(6)))
        for c in range(2):
            lines.append(fmt(win.T_s_f[c]) + " " + fmt(win.K[c]))
        pos += 1
    for k, f in enumerate(order):             # time order
        frame_time_pos[f] = pos
        ts = 1_000_000_000 + k * 250_000_000
        has_imu = 1 if win.vio else 0
        lines.append(f"{ts} {0 if f in nonkf else 1} 1 {int(win.has_prior[f])} {has_imu} 2")
        lines.append(fmt(win.T_f_w[f]) + " " + fmt(win.T_prior[f]) + " " + fmt(win.inf_prior[f]))
        if has_imu:
            ps = [p for p in range(win.n_imu) if win.imu_j[p] == f]
            if ps:
                p = ps[0]
                pre = [win.imu_dR[p], win.imu_dv[p], win.imu_dp[p], win.imu_cov[p], win.imu_J_dR_bg[p], win.imu_J_dv_ba[p],
                       win.imu_J_dv_bg[p], win.imu_J_dp_ba[p], win.imu_J_dp_bg[p]]
                lastkf = frame_time_pos[int(win.imu_i[p])]
            else:
                pre = [np.eye(3), np.zeros(3), np.zeros(3), np.zeros(81)] + [np.zeros(9)] * 5
                lastkf = -1
            lines.append(fmt(win.v[f]) + " " + fmt(win.ba[f]) + " " + fmt(win.bg[f]) + " " + " ".join(fmt(x) for x in pre) +
                         f" {float(synth.BACC_NOISE)!r} {float(synth.BGYR_NOISE)!r} {lastkf}")
        for c in range(2):
            lines.append(fmt(win.T_s_f[c]) + " " + fmt(win.K[c]))
        pos += 1
    # landmarks
    L = win.n_lmks
    ptr = np.searchsorted(win.obs_lmk, np.arange(L + 1))
    lm_lines = []
    exp_triplets, exp_lmk_of = [], []
    n_kept = 0
    for l in range(L):
        outlier = extras and l % 7 == 3
        uninit = extras and l % 11 == 5
        feats = []
        for o in range(ptr[l], ptr[l + 1]):
            feats.append((frame_time_pos[int(win.obs_frame[o])], int(win.obs_cam[o]), win.obs_bearing[o], win.obs_uv[o], 1,
                          int(win.obs_frame[o])))
        if extras and l % 5 == 0:
            feats.insert(1, (0, 0, np.array([0.0, 0.0, 1.0]), np.zeros(2), 1, None))       # feature in a frame outside the window
        if extras and l % 5 == 1:
            feats.insert(0, (feats[0][0], feats[0][1], np.array([0.0, 0.0, 1.0]), np.zeros(2), 0, None))  # expired weak_ptr
        lm_lines.append(fmt(win.lmk_t[l]) + f" {0 if uninit else 1} {1 if outlier else 0} {len(feats)}")
        for (fp, c, b, uv, alive, f_win) in feats:
            lm_lines.append(f"{fp} {c} " + fmt(b) + " " + fmt(uv) + f" {alive}")
        if outlier or uninit:
            continue
        for (fp, c, b, uv, alive, f_win) in feats:
            if not alive or f_win is None or f_win in nonkf:
                continue
            exp_triplets.append((n_kept, f_win, c))
        n_kept += 1
    lines.append(str(L))
    lines += lm_lines
    exp_imu = []
    if win.vio:
        for f in range(F):                    # frame_vector order of j (newest first)
            ps = [p for p in range(win.n_imu) if win.imu_j[p] == f]
            if ps:
                exp_imu.append((int(win.imu_i[ps[0]]), f))
    return "\n".join(lines) + "\n", np.array(exp_triplets, dtype=np.int64).reshape(-1, 3), exp_imu, n_kept


@pytest.mark.parametrize("extras", [False, True])
def test_flatten_order_is_bit_exact(adapter_exe, extras):
    win = synth.make_window("small")
    txt, exp, exp_imu, n_kept = graph_text(win, np.random.default_rng(0), extras)
    out = subprocess.run([adapter_exe, "flatten", "1", "1"], input=txt, capture_output=True, text=True, check=True).stdout.split("\n")
    assert out[0] == "flatten"
    F, C, L, O, P = (int(x) for x in out[1].split())
    assert (F, C, L) == (win.n_frames, 2, n_kept)
    got = np.array([[int(x) for x in ln.split()] for ln in out[2:2 + O]], dtype=np.int64).reshape(-1, 3)
    assert O == exp.shape[0]
    assert np.array_equal(got, exp)                       # visibility triplets, bit-exact, in walk order
    if not extras:
        assert np.array_equal(got[:, 0], win.obs_lmk) and np.array_equal(got[:, 1], win.obs_frame) and np.array_equal(got[:, 2], win.obs_cam)
    imu = [tuple(int(x) for x in ln.split()[:2]) for ln in out[2 + O:2 + O + P]]
    assert imu == exp_imu
    tx = np.array([float(ln) for ln in out[2 + O + P:2 + O + P + F]])
    assert np.array_equal(tx, win.T_f_w[:, 3])            # frames newest -> oldest


@pytest.mark.gpu
def test_adapter_solve_matches_python_binding(adapter_exe):
    from sadvio_b200 import api

    win = synth.make_window("small")
    txt, _, _, _ = graph_text(win, np.random.default_rng(0), False)
    out = subprocess.run([adapter_exe, "solve", "1", "1"], input=txt, capture_output=True, text=True, check=True).stdout.split("\n")
    ok, iters = (int(x) for x in out[1].split())
    assert ok == 1
    opt = api.B200Optimizer()
    ref = synth.make_window("small")
    assert opt.localMapVIOptimization(ref, 1)
    assert iters == opt.last_stats["iterations"]
    F = win.n_frames
    rows = [np.array([float(x) for x in ln.split()]) for ln in out[2:2 + F]]   # oldest -> newest
    for k, row in enumerate(rows):
        f = F - 1 - k
        assert np.abs(row[:12] - ref.T_f_w[f]).max() < 1e-9
        assert np.abs(row[12:15] - ref.v[f]).max() < 1e-9
        assert np.abs(row[15:18] - ref.ba[f]).max() < 1e-9
        ps = [p for p in range(ref.n_imu) if ref.imu_j[p] == f]
        if ps:
            assert np.abs(row[21:24] - ref.imu_dp[ps[0]]).max() < 1e-9   # biasDeltaCorrection applied
    lm = np.array([[float(x) for x in ln.split()] for ln in out[2 + F:2 + F + win.n_lmks]])
    assert np.abs(lm - ref.lmk_t).max() < 1e-8

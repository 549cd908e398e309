"""Python binding of the C ABI (include/sdv.h) and a host-side mirror of the reference optimizer interface.

The product path is the CUDA library only: if ``libsadvio_b200.so`` is missing or no sm_100 GPU is visible, calls fail
loudly (``BackendUnavailable``); there is no CPU fallback and this module never imports ``oracle``.

``B200Optimizer`` mirrors ``isae::AOptimizer`` for the entry points this repository replaces
(reference cpp/include/isaeslam/optimizers/AOptimizer.h:22-30):
    bool localMapBA(local_map, fixed_frame_number = 0)
    bool localMapVIOptimization(local_map, fixed_frame_number = 0)
    bool landmarkOptimization(frame) / singleFrameOptimization(frame) / singleFrameVIOptimization(frame)
operating on a flattened ``abi.Window`` (the C++ adapter sadvio_b200/host/b200_optimizer.hpp does the flattening of the
pointer graph; here a frame is an index into the window).  All return ``True``/``False`` like the reference and update
the window state in place exactly as AOptimizer.cpp:122-146, :199-216, :262-296, :391-434 do.
"""
from __future__ import annotations

import atexit
import ctypes as C
import os
import weakref

import numpy as np

from . import abi, build as _build
from .synth import exp_so3


class BackendUnavailable(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path):
        raise BackendUnavailable(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(path)
    dp = abi.c_double_p
    L.sdv_abi_version.restype = C.c_int
    L.sdv_strerror.restype = C.c_char_p
    L.sdv_strerror.argtypes = [C.c_int]
    L.sdv_last_error.restype = C.c_char_p
    L.sdv_last_error.argtypes = [C.c_void_p]
    L.sdv_default_config.argtypes = [C.POINTER(abi.SdvConfig)]
    L.sdv_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(abi.SdvConfig)]
    L.sdv_destroy.argtypes = [C.c_void_p]
    L.sdv_solve_window.argtypes = [C.c_void_p, C.POINTER(abi.SdvWindow), C.POINTER(abi.SdvDelta), C.POINTER(abi.SdvStats)]
    L.sdv_upload_window.argtypes = [C.c_void_p, C.POINTER(abi.SdvWindow)]
    L.sdv_solve_resident.argtypes = [C.c_void_p, C.POINTER(abi.SdvStats)]
    L.sdv_download_delta.argtypes = [C.c_void_p, C.POINTER(abi.SdvDelta)]
    L.sdv_eval_visual.argtypes = [C.c_void_p, C.POINTER(abi.SdvDelta), dp, dp, dp, dp]
    L.sdv_eval_imu.argtypes = [C.c_void_p, C.POINTER(abi.SdvDelta), dp, dp, dp]
    L.sdv_comm_unique_id.argtypes = [C.c_void_p]
    L.sdv_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
    L.sdv_comm_peer_handle.argtypes = [C.c_void_p, C.c_void_p]
    L.sdv_comm_peer_open.argtypes = [C.c_void_p, C.c_void_p]
    L.sdv_time_kernel.argtypes = [C.c_void_p, C.c_int32, C.c_int32, dp]
    L.sdv_debug_read.argtypes = [C.c_void_p, C.c_int32, dp, C.c_int64]
    L.sdv_debug_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.sdv_debug_graph_builds.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.sdv_preintegrate.argtypes = [C.c_void_p, C.POINTER(abi.SdvImuIntervals), C.POINTER(abi.SdvPreint)]
    L.sdv_marginalize.argtypes = [C.c_void_p, C.POINTER(abi.SdvWindow), C.c_int32, C.POINTER(abi.SdvMarginalSizes)]
    L.sdv_marginal_fetch.argtypes = [C.c_void_p, C.POINTER(abi.SdvMarginal)]
    L.sdv_viinit.argtypes = [C.c_void_p, C.POINTER(abi.SdvWindow), C.c_int32, C.POINTER(abi.SdvViinitResult), C.POINTER(abi.SdvStats)]
    L.sdv_schur_prior.argtypes = [C.c_void_p, dp, dp, C.c_int32, C.c_int32, C.c_double, C.POINTER(abi.SdvMarginalSizes)]
    _lib = L
    return L


def default_config() -> abi.SdvConfig:
    cfg = abi.SdvConfig()
    lib().sdv_default_config(C.byref(cfg))
    return cfg


_live_solvers: "weakref.WeakSet[Solver]" = weakref.WeakSet()


def _close_all_solvers():
    # at interpreter exit, BEFORE torch / NCCL tear their own state down (atexit runs last-registered first): a handle with a
    # communicator that is destroyed later, from __del__ during shutdown, was seen to stall multi-rank runs for minutes
    for s in list(_live_solvers):
        try:
            s.close()
        except Exception:  # noqa: BLE001
            pass


class Solver:
    """Thin RAII wrapper around ``sdv_handle``."""

    def __init__(self, cfg: abi.SdvConfig | None = None, device: int = 0):
        L = lib()
        if not _live_solvers:
            atexit.unregister(_close_all_solvers)
            atexit.register(_close_all_solvers)
        self.cfg = cfg or default_config()
        self.cfg.device = device
        self._h = C.c_void_p()
        rc = L.sdv_create(C.byref(self._h), C.byref(self.cfg))
        if rc != 0:
            self._h = None
            raise BackendUnavailable(f"sdv_create failed: {L.sdv_strerror(rc).decode()}")
        self._win = None
        _live_solvers.add(self)

    def close(self):
        if getattr(self, "_h", None):
            lib().sdv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, allow=(0,)):
        if rc not in allow:
            L = lib()
            raise RuntimeError(f"sdv error {rc} ({L.sdv_strerror(rc).decode()}): {L.sdv_last_error(self._h).decode()}")
        return rc

    # -- the reference-facing call: host buffers in, host buffers out
    def solve_window(self, win: abi.Window):
        ws = win.as_struct()
        out = abi.Delta.zeros(win.n_frames, win.n_lmks)
        ds = out.as_struct()
        st = abi.SdvStats()
        rc = self._check(lib().sdv_solve_window(self._h, C.byref(ws), C.byref(ds), C.byref(st)), allow=(0, 5))
        self._win = win
        return rc, out, abi.stats_to_dict(st)

    # -- resident variant (bench: inputs already in HBM)
    def upload(self, win: abi.Window):
        ws = win.as_struct()
        self._check(lib().sdv_upload_window(self._h, C.byref(ws)))
        self._win = win

    def solve_resident(self):
        st = abi.SdvStats()
        rc = self._check(lib().sdv_solve_resident(self._h, C.byref(st)), allow=(0, 5))
        return rc, abi.stats_to_dict(st)

    def download(self) -> abi.Delta:
        out = abi.Delta.zeros(self._win.n_frames, self._win.n_lmks)
        ds = out.as_struct()
        self._check(lib().sdv_download_delta(self._h, C.byref(ds)))
        return out

    def eval_visual(self, x: abi.Delta | None = None):
        O = self._win.n_obs
        r, Jp, Jl, cost = np.zeros((O, 2)), np.zeros((O, 12)), np.zeros((O, 6)), np.zeros(1)
        xs = x.as_struct() if x is not None else None
        p = lambda a: a.ctypes.data_as(abi.c_double_p)
        self._check(lib().sdv_eval_visual(self._h, C.byref(xs) if xs is not None else None, p(r), p(Jp), p(Jl), p(cost)))
        return r, Jp, Jl, float(cost[0])

    def eval_imu(self, x: abi.Delta | None = None):
        P = self._win.n_imu
        r, J, rb = np.zeros((P, 9)), np.zeros((P, 216)), np.zeros((P, 6))
        xs = x.as_struct() if x is not None else None
        p = lambda a: a.ctypes.data_as(abi.c_double_p)
        self._check(lib().sdv_eval_imu(self._h, C.byref(xs) if xs is not None else None, p(r), p(J), p(rb)))
        return r, J.reshape(P, 9, 24), rb

    def time_kernel(self, which: int, repeats: int = 20) -> float:
        ms = np.zeros(1)
        self._check(lib().sdv_time_kernel(self._h, which, repeats, ms.ctypes.data_as(abi.c_double_p)))
        return float(ms[0])

    def debug_dims(self):
        n, npad = C.c_int32(), C.c_int32()
        self._check(lib().sdv_debug_dims(self._h, C.byref(n), C.byref(npad)))
        return n.value, npad.value

    def preintegrate(self, sample_ptr, acc, gyr, dt, T_f_w, v, ba, bg, eta, rate_hz, dR_stale=None) -> dict:
        """IMU::processIMU (IMU.cpp:5-91) over the samples of every keyframe interval, on the GPU (sdv_preintegrate)."""
        f64 = lambda a, shape: np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(shape))
        sp = np.ascontiguousarray(sample_ptr, dtype=np.int32)
        n, S = len(sp) - 1, int(sp[-1])
        ins = abi.SdvImuIntervals()
        ins.n_intervals, ins.n_samples = n, S
        keep = [sp, f64(acc, (S, 3)), f64(gyr, (S, 3)), f64(dt, (S,)), f64(T_f_w, (n, 12)), f64(v, (n, 3)), f64(ba, (n, 3)), f64(bg, (n, 3))]
        ins.sample_ptr = keep[0].ctypes.data_as(abi.c_int32_p)
        for name, a in zip(("acc", "gyr", "dt", "T_f_w", "v", "ba", "bg"), keep[1:]):
            setattr(ins, name, a.ctypes.data_as(abi.c_double_p))
        if dR_stale is not None:
            keep.append(f64(dR_stale, (n, 9)))
            ins.dR_stale = keep[-1].ctypes.data_as(abi.c_double_p)
        ins.eta[:] = [float(x) for x in np.asarray(eta).reshape(6)]
        ins.rate_hz = float(rate_hz)
        out = {k: np.zeros((n, w)) for k, w in (("dR", 9), ("dv", 3), ("dp", 3), ("cov", 81), ("J_dR_bg", 9), ("J_dv_ba", 9), ("J_dv_bg", 9),
                                              ("J_dp_ba", 9), ("J_dp_bg", 9), ("T_pred", 12), ("v_pred", 3))}
        outs = abi.SdvPreint()
        for k, a in out.items():
            setattr(outs, k, a.ctypes.data_as(abi.c_double_p))
        self._check(lib().sdv_preintegrate(self._h, C.byref(ins), C.byref(outs)))
        return out

    # -- marginal-prior construction (AngularAdjustmentCERESAnalytic::marginalize and what it calls, on the GPU)
    def _fetch_marginal(self, sz, with_A: bool):
        n, nf, N, K = sz.n, sz.n_full, sz.m + sz.n, sz.n_keep
        out = dict(J=np.zeros((nf, n)), r0=np.zeros(nf), keep=np.zeros(K, dtype=np.int32), marg=np.zeros(sz.n_marg, dtype=np.int32), Ak=np.zeros((n, n)),
                   bk=np.zeros(n), U=np.zeros((n, nf)), Lambda=np.zeros(nf), imu_sqrt_inf=np.zeros(225), p2l_delta=np.zeros((K, 3)),
                   p2l_sqrt_inf=np.zeros((K, 9)), chain=np.zeros(max(sz.n_chain, 1), dtype=np.int32), l2l_delta=np.zeros((max(sz.n_chain - 1, 1), 3)),
                   l2l_sqrt_inf=np.zeros((max(sz.n_chain - 1, 1), 9)))
        if with_A:
            out.update(A=np.zeros((N, N)), b=np.zeros(N))
        ms = abi.SdvMarginal()
        for k, name in (("J", "J"), ("r0", "r0"), ("Ak", "Ak"), ("bk", "bk"), ("U", "U"), ("Lambda", "Lambda"), ("A", "A"), ("b", "b"),
                        ("imu_sqrt_inf", "imu_sqrt_inf"), ("p2l_delta", "p2l_delta"), ("p2l_sqrt_inf", "p2l_sqrt_inf"), ("l2l_delta", "l2l_delta"),
                        ("l2l_sqrt_inf", "l2l_sqrt_inf")):
            if name in out:
                setattr(ms, k, out[name].ctypes.data_as(abi.c_double_p))
        ms.keep_lmk = out["keep"].ctypes.data_as(abi.c_int32_p)
        ms.marg_lmk = out["marg"].ctypes.data_as(abi.c_int32_p)
        ms.chain = out["chain"].ctypes.data_as(abi.c_int32_p)
        self._check(lib().sdv_marginal_fetch(self._h, C.byref(ms)))
        out["lmk_with_prior"] = int(ms.lmk_with_prior)
        out["lmk_sqrt_inf"] = np.array(ms.lmk_sqrt_inf[:])
        out["chain"] = out["chain"][:sz.n_chain]
        out["l2l_delta"], out["l2l_sqrt_inf"] = out["l2l_delta"][:max(sz.n_chain - 1, 0)], out["l2l_sqrt_inf"][:max(sz.n_chain - 1, 0)]
        return out

    def marginalize(self, win: abi.Window, sparsify: bool = False):
        """marginalize(frame0 = oldest keyframe of `win`, frame1 = the next one, enable_sparsif) on the GPU.  Returns
        (dense, sparse, info): abi.DensePrior over (frame 1, kept landmarks) in the landmark indices of `win` — None when the
        reference's marginalize() returns false —, the abi.SparsePrior of sparsifyVIO / sparsifyVO when asked for, and a dict
        with the intermediate results (A, b, Ak, bk, U, Lambda, keep, marg, sizes)."""
        ws = win.as_struct()
        sz = abi.SdvMarginalSizes()
        self._check(lib().sdv_marginalize(self._h, C.byref(ws), 1 if sparsify else 0, C.byref(sz)))
        self._win = win
        info = {k: getattr(sz, k) for k, _ in abi.SdvMarginalSizes._fields_}
        if not sz.ok:
            return None, None, info
        out = self._fetch_marginal(sz, True)
        info.update(out)
        first = 15 if sz.frame >= 0 else 0
        dense = abi.DensePrior(J=out["J"], r0=out["r0"], frame=int(sz.frame), frame_col=0, keep_lmk=out["keep"].copy(),
                               keep_col=np.asarray([first + 3 * k for k in range(sz.n_keep)], dtype=np.int32))
        sparse = None
        if sparsify and sz.n_full > 0:
            f1 = int(sz.frame)
            if f1 >= 0:
                sparse = abi.SparsePrior(has_imu_prior=True, frame=f1, T_prior=win.T_f_w[f1].copy(), v_prior=win.v[f1].copy(), ba_prior=win.ba[f1].copy(),
                                         bg_prior=win.bg[f1].copy(), imu_sqrt_inf=out["imu_sqrt_inf"], p2l_lmk=out["keep"].copy(),
                                         p2l_delta=out["p2l_delta"], p2l_sqrt_inf=out["p2l_sqrt_inf"])
            elif sz.n_chain >= 2:
                wp = out["lmk_with_prior"]
                sparse = abi.SparsePrior(has_lmk_prior=True, lmk0=wp, lmk_prior=win.lmk_t[wp].copy(), lmk_sqrt_inf=out["lmk_sqrt_inf"],
                                         l2l_a=out["chain"][:-1].copy(), l2l_b=out["chain"][1:].copy(), l2l_delta=out["l2l_delta"],
                                         l2l_sqrt_inf=out["l2l_sqrt_inf"])
        return dense, sparse, info

    def schur_prior(self, A, b, m: int, eps: float = 1e-12):
        """computeSchurComplement + rankReveallingDecomposition + computeJacobiansAndResiduals on a given information matrix."""
        A = np.ascontiguousarray(A, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        n = A.shape[0] - m
        sz = abi.SdvMarginalSizes()
        self._check(lib().sdv_schur_prior(self._h, A.ctypes.data_as(abi.c_double_p), b.ctypes.data_as(abi.c_double_p), m, n, eps, C.byref(sz)))
        if not sz.ok:
            return None
        out = self._fetch_marginal(sz, False)
        out.update({k: getattr(sz, k) for k, _ in abi.SdvMarginalSizes._fields_})
        return out

    def viinit(self, win: abi.Window, optim_scale: bool = True):
        """sdv_viinit: the solve of AOptimizer::VIInit over the frames / IMU pairs of `win`.
        Returns rc, dict(dv[F][3], r_wi[2], lam, R_w_i[3][3], scale), stats."""
        ws = win.as_struct()
        dv = np.zeros((win.n_frames, 3))
        res = abi.SdvViinitResult()
        res.dv = dv.ctypes.data_as(abi.c_double_p)
        st = abi.SdvStats()
        rc = self._check(lib().sdv_viinit(self._h, C.byref(ws), int(bool(optim_scale)), C.byref(res), C.byref(st)), allow=(0, 5))
        return rc, dict(dv=dv, r_wi=np.array(res.r_wi[:]), lam=float(res.lambda_), R_w_i=np.array(res.R_w_i[:]).reshape(3, 3),
                        scale=float(res.scale)), abi.stats_to_dict(st)

    def graph_builds(self) -> int:
        n = C.c_int64()
        self._check(lib().sdv_debug_graph_builds(self._h, C.byref(n)))
        return int(n.value)

    def debug_read(self, what: int, count: int) -> np.ndarray:
        out = np.zeros(count)
        self._check(lib().sdv_debug_read(self._h, what, out.ctypes.data_as(abi.c_double_p), count))
        return out

    def comm_init(self, uid: bytes, rank: int, world: int):
        buf = C.create_string_buffer(uid, 128)
        self._check(lib().sdv_comm_init(self._h, buf, rank, world))

    def comm_peer_handle(self) -> bytes:
        """64-byte CUDA IPC handle of this rank's exchange area (sdv_comm_peer_handle)."""
        buf = C.create_string_buffer(64)
        self._check(lib().sdv_comm_peer_handle(self._h, buf))
        return buf.raw

    def comm_peer_open(self, handles: bytes):
        """`handles`: the 64-byte handles of all ranks, concatenated in rank order."""
        buf = C.create_string_buffer(handles, len(handles))
        self._check(lib().sdv_comm_peer_open(self._h, buf))


def comm_setup(solver: "Solver", dist, device) -> None:
    """The whole multi-GPU setup of a handle through torch.distributed (backend nccl): the NCCL unique id from rank 0, then the
    all-gather of the CUDA IPC handles of the peer-memory exchange areas.  `SDV_NO_PEER=1` keeps the exchanges on NCCL."""
    import torch

    rank, world = dist.get_rank(), dist.get_world_size()
    uid = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        uid = torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8).to(device)
    dist.broadcast(uid, 0)
    solver.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
    if world <= 8 and not os.environ.get("SDV_NO_PEER_SETUP"):
        mine = torch.frombuffer(bytearray(solver.comm_peer_handle()), dtype=torch.uint8).to(device)
        allh = [torch.zeros(64, dtype=torch.uint8, device=device) for _ in range(world)]
        dist.all_gather(allh, mine)
        solver.comm_peer_open(b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh))


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = lib().sdv_comm_unique_id(buf)
    if rc != 0:
        raise RuntimeError("sdv_comm_unique_id failed")
    return buf.raw


def write_back(win: abi.Window, d: abi.Delta, vio: bool) -> None:
    """State write-back of the window solves, in place (AOptimizer.cpp:391-434 for VIO, :328-341 for BA)."""
    for f in range(win.n_frames):
        T = np.vstack([win.T_f_w[f].reshape(3, 4), [0, 0, 0, 1]])
        dT = np.eye(4)
        dT[:3, :3] = exp_so3(d.dpose[f, :3])
        dT[:3, 3] = d.dpose[f, 3:]
        win.T_f_w[f] = (T @ dT)[:3, :4].reshape(12)
    win.lmk_t += d.dlmk
    if vio:
        win.v += d.dv
        win.ba += d.dba
        win.bg += d.dbg
        # IMU::biasDeltaCorrection with the PREVIOUS keyframe's dba, dbg (AOptimizer.cpp:421-434, IMU.cpp:104-108) for EVERY
        # frame whose getLastKF() owns dba / dbg blocks: the frames with an IMUFactor (their previous keyframe is imu_i) and
        # the ones the factor loop skipped (dt > 1 s), which the window carries host-side in `skipped_preint`
        def correct(dR, dv, dp, J_dR_bg, J_dv_ba, J_dv_bg, J_dp_ba, J_dp_bg, k, i):
            dba, dbg = d.dba[i], d.dbg[i]
            dp[k] += J_dp_ba[k].reshape(3, 3) @ dba + J_dp_bg[k].reshape(3, 3) @ dbg
            dv[k] += J_dv_ba[k].reshape(3, 3) @ dba + J_dv_bg[k].reshape(3, 3) @ dbg
            dR[k] = (dR[k].reshape(3, 3) @ exp_so3(J_dR_bg[k].reshape(3, 3) @ dbg)).reshape(9)

        for p in range(win.n_imu):
            correct(win.imu_dR, win.imu_dv, win.imu_dp, win.imu_J_dR_bg, win.imu_J_dv_ba, win.imu_J_dv_bg, win.imu_J_dp_ba,
                    win.imu_J_dp_bg, p, int(win.imu_i[p]))
        sk = win.skipped_preint
        if sk is not None:
            for k in range(len(sk.frame)):
                correct(sk.dR, sk.dv, sk.dp, sk.J_dR_bg, sk.J_dv_ba, sk.J_dv_bg, sk.J_dp_ba, sk.J_dp_bg, k, int(sk.prev[k]))


def viinit_write_back(win: abi.Window, res: dict) -> None:
    """State update of AOptimizer::VIInit, in place (AOptimizer.cpp:531-567): velocities, poses rotated into the inertial frame
    and re-scaled, priors re-set on the new poses, landmarks transformed.  (The reference adds dba — a constant block, zero —
    to the accelerometer bias twice and never touches the gyroscope bias, :535-536: a no-op.)"""
    win.v += res["dv"]                                                     # :533-534
    R_w_i, s = res["R_w_i"], float(np.exp(res["lam"]))                     # :540
    for f in range(win.n_frames):                                          # :545-556
        T = win.T_f_w[f].reshape(3, 4).copy()
        T[:, 3] *= s                                                       # :549
        T[:, :3] = T[:, :3] @ R_w_i                                        # :550 (T_w_i has no translation)
        win.T_f_w[f] = T.reshape(12)
        if win.has_prior is not None and win.has_prior[f]:                 # :553-555
            win.T_prior[f] = win.T_f_w[f]
            win.inf_prior[f] = 100.0
    if win.n_lmks:                                                         # :559-567
        win.lmk_t[:] = s * (win.lmk_t @ R_w_i)                             # exp(lambda) R_w_i^T t, row-wise


HUBER_A = float(np.sqrt(1.345))  # AOptimizer.cpp:102, :223


def _subset(win: abi.Window, frames: list[int], lmks: np.ndarray, obs: np.ndarray, **kw) -> abi.Window:
    """The window restricted to `frames` (new order), `lmks` and the observations `obs` (landmark-major already)."""
    fmap = -np.ones(win.n_frames, dtype=np.int64)
    fmap[frames] = np.arange(len(frames))
    lmap = -np.ones(win.n_lmks, dtype=np.int64)
    lmap[lmks] = np.arange(len(lmks))
    pick = lambda a, idx: None if a is None else np.ascontiguousarray(a[idx])
    return abi.Window(
        vio=False, factor_kind=win.factor_kind, n_fixed=0, T_f_w=pick(win.T_f_w, frames), T_s_f=win.T_s_f.copy(), K=win.K.copy(),
        lmk_t=pick(win.lmk_t, lmks), obs_lmk=lmap[win.obs_lmk[obs]].astype(np.int32), obs_frame=fmap[win.obs_frame[obs]].astype(np.int32),
        obs_cam=pick(win.obs_cam, obs), obs_bearing=pick(win.obs_bearing, obs), obs_uv=pick(win.obs_uv, obs),
        v=pick(win.v, frames), ba=pick(win.ba, frames), bg=pick(win.bg, frames), has_imu=pick(win.has_imu, frames), **kw)


def landmark_window(win: abi.Window, frame: int) -> tuple[abi.Window, np.ndarray]:
    """landmarkOptimization(frame) as a window (addLandmarkResiduals, AngularAdjustmentCERESAnalytic.cpp:106-209): the
    landmarks `frame` observes, all their observations on the keyframes of `win`, every pose constant, Huber loss, 10
    iterations.  Returns the window and the indices of its landmarks in `win`."""
    lmks = np.unique(win.obs_lmk[win.obs_frame == frame])
    obs = np.flatnonzero(np.isin(win.obs_lmk, lmks))
    frames = list(dict.fromkeys(int(f) for f in win.obs_frame[obs]))  # first-appearance order
    sub = _subset(win, frames, lmks, obs, visual_loss_huber_a=HUBER_A, max_num_iterations=10)
    sub.n_fixed = len(frames)
    return sub, lmks


def single_frame_window(win: abi.Window, frame: int, vi: bool) -> tuple[abi.Window, list[int]]:
    """singleFrameOptimization / singleFrameVIOptimization(frame) as a window (addSingleFrameResiduals,
    AngularAdjustmentCERESAnalytic.cpp:6-102; AOptimizer.cpp:152-297): the frame (and, VI, its previous keyframe = the i of
    its IMU pair) with free poses, the frame's landmarks constant, sigma 1 / focal, 5 iterations."""
    pair = np.flatnonzero(win.imu_j == frame) if (vi and win.vio and win.imu_j is not None) else np.zeros(0, dtype=np.int64)
    frames = [frame] + ([int(win.imu_i[pair[0]])] if len(pair) else [])
    lmks = np.unique(win.obs_lmk[win.obs_frame == frame])
    sel = np.isin(win.obs_lmk, lmks) & np.isin(win.obs_frame, frames)
    # landmark-major; within a landmark the moving frame's features first, then the previous keyframe's
    obs = np.flatnonzero(sel)
    obs = obs[np.lexsort((obs, win.obs_frame[obs] != frame, win.obs_lmk[obs]))]
    sub = _subset(win, frames, lmks, obs, landmarks_constant=True, max_num_iterations=5, visual_loss_huber_a=HUBER_A if vi else 0.0)
    if win.factor_kind == abi.SDV_FACTOR_ANGULAR:
        focal = (win.K[:, 0] + win.K[:, 1]) / 2
        sub.obs_sigma = np.ascontiguousarray(1.0 / focal[sub.obs_cam])  # …Analytic.cpp:48
    if len(pair):
        p = pair[:1]
        sub.vio = True
        sub.imu_i, sub.imu_j = np.array([1], dtype=np.int32), np.array([0], dtype=np.int32)
        for name in ("imu_dt", "imu_dR", "imu_dv", "imu_dp", "imu_cov", "imu_J_dR_bg", "imu_J_dv_ba", "imu_J_dv_bg", "imu_J_dp_ba", "imu_J_dp_bg",
                     "imu_sigma_ba", "imu_sigma_bg"):
            setattr(sub, name, np.ascontiguousarray(getattr(win, name)[p]))
    return sub, frames


class B200Optimizer:
    """Host-side mirror of ``isae::AOptimizer`` for the window solves (see module docstring)."""

    def __init__(self, device: int = 0, cfg: abi.SdvConfig | None = None):
        self.solver = Solver(cfg, device)
        self.last_stats: dict | None = None
        self.marginalization = (None, None)
        self.last_marginal_info: dict | None = None

    def _solve(self, win: abi.Window):
        return self.solver.solve_window(win)

    def _run(self, win: abi.Window, fixed_frame_number: int, vio: bool) -> bool:
        win.n_fixed = int(fixed_frame_number)
        win.vio = vio
        try:
            rc, d, st = self._solve(win)  # rc 5 (Ceres FAILURE) still carries the last accepted x
        except RuntimeError:
            return False  # no solve ran (malformed window, CUDA error): state untouched, like the C++ adapter
        self.last_stats = st
        write_back(win, d, vio)  # the reference ignores the summary: write back, return true (AOptimizer.cpp:388-445)
        return True

    def localMapVIOptimization(self, local_map: abi.Window, fixed_frame_number: int = 0) -> bool:  # noqa: N802
        return self._run(local_map, fixed_frame_number, True)

    def localMapBA(self, local_map: abi.Window, fixed_frame_number: int = 0) -> bool:  # noqa: N802
        return self._run(local_map, fixed_frame_number, False)

    def marginalize(self, local_map: abi.Window, enable_sparsif: bool = False) -> bool:
        """AOptimizer::marginalize(frame0 = the oldest keyframe of the window, frame1 = the next one, enable_sparsif)
        (AngularAdjustmentCERESAnalytic.cpp:488-739) on the GPU.  The result stays on the optimizer like `_marginalization`:
        `self.marginalization = (dense, sparse)` in the landmark indices of `local_map`; `local_map.dense_prior` is read as
        `_marginalization_last`.  False (and the scheme reset) when the reference's marginalize() returns false."""
        try:
            dense, sparse, info = self.solver.marginalize(local_map, enable_sparsif)
        except RuntimeError:
            return False
        self.last_marginal_info = info
        self.marginalization = (dense, sparse)
        return dense is not None

    def VIInit(self, local_map: abi.Window, optim_scale: bool = False):  # noqa: N802
        """AOptimizer::VIInit(local_map, R_w_i, optim_scale) (AOptimizer.cpp:448-581): returns (exp(lambda), R_w_i) and updates
        the map in place.  `local_map.imu_*` must list every (getLastKF(), frame) pair — VIInit has no dt test (:485-502)."""
        try:
            rc, res, st = self.solver.viinit(local_map, optim_scale)
        except RuntimeError:
            return float("nan"), None   # no solve ran (malformed input, CUDA error): state untouched, like the C++ adapter
        self.last_stats = st
        viinit_write_back(local_map, res)   # the reference ignores the summary here too
        return res["scale"], res["R_w_i"]

    def landmarkOptimization(self, local_map: abi.Window, frame: int, sanity_check=None) -> bool:  # noqa: N802
        """AOptimizer.cpp:98-150.  `sanity_check(l) -> bool` stands for ALandmark::sanityCheck (a data-model method outside the
        optimizer): only landmarks that pass it are updated (:132-140); None = all pass."""
        sub, lmks = landmark_window(local_map, frame)
        if sub.n_obs == 0:
            return True
        try:
            rc, d, st = self._solve(sub)
        except RuntimeError:
            return False
        self.last_stats = st
        for k, l in enumerate(lmks):
            if sanity_check is None or sanity_check(int(l)):
                local_map.lmk_t[l] += d.dlmk[k]
        return True

    def _single(self, local_map: abi.Window, frame: int, vi: bool) -> bool:
        sub, frames = single_frame_window(local_map, frame, vi)
        if sub.n_obs == 0 and sub.n_imu == 0:
            return True
        try:
            rc, d, st = self._solve(sub)
        except RuntimeError:
            return False
        self.last_stats = st
        if vi and rc == 5:
            return False  # !summary.IsSolutionUsable(), AOptimizer.cpp:259
        sub.lmk_t = np.zeros((0, 3))
        d.dlmk = np.zeros((0, 3))
        sub.imu_i = sub.imu_j = None  # no biasDeltaCorrection in the single-frame solves (AOptimizer.cpp:262-285)
        write_back(sub, d, bool(sub.vio))
        local_map.T_f_w[frames] = sub.T_f_w
        if sub.vio:
            local_map.v[frames], local_map.ba[frames], local_map.bg[frames] = sub.v, sub.ba, sub.bg
        return True

    def singleFrameOptimization(self, local_map: abi.Window, frame: int) -> bool:  # noqa: N802
        return self._single(local_map, frame, False)

    def singleFrameVIOptimization(self, local_map: abi.Window, frame: int) -> bool:  # noqa: N802
        return self._single(local_map, frame, True)

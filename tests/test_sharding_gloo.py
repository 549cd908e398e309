"""N > 1 host logic on CPU: world_size-2 (and 3) gloo processes, each building the reduced system of ITS landmark shard
(sdv_shard_range = the partition sdv_upload_window applies) with the oracle; a gloo all-reduce(sum) must reproduce the
full reduced system — the exchange the NCCL path performs once per LM iteration."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard_range(win, rank, world):
    from sadvio_b200 import api

    L = api.lib()
    L.sdv_shard_range.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_int32, C.c_int32] + [C.POINTER(C.c_int32)] * 4
    out = [C.c_int32() for _ in range(4)]
    rc = L.sdv_shard_range(win.obs_lmk.ctypes.data_as(C.POINTER(C.c_int32)), win.n_obs, win.n_lmks, rank, world, *[C.byref(o) for o in out])
    assert rc == 0
    return tuple(o.value for o in out)


def _worker(rank, world, port, name, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    from sadvio_b200 import synth

    win = synth.make_window(name)
    l0, l1, o0, o1 = shard_range(win, rank, world)
    # non-visual factors (IMU, priors) are assembled on rank 0 only
    S, g = oracle.reduced_system(win, l0, l1, with_factors=(rank == 0))
    buf = torch.from_numpy(np.concatenate([S.reshape(-1), g, [float(o1 - o0), float(l1 - l0)]]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)  # ONE all-reduce of [S | g | ...]
    ranges = [None] * world
    dist.all_gather_object(ranges, (l0, l1, o0, o1))
    if rank == 0:
        S_full, g_full = oracle.reduced_system(win, 0, win.n_lmks, True)
        n = S_full.shape[0]
        Sr = buf[: n * n].numpy().reshape(n, n)
        gr = buf[n * n: n * n + n].numpy()
        q.put(dict(err_S=float(np.abs(Sr - S_full).max() / np.abs(S_full).max()), err_g=float(np.abs(gr - g_full).max() / np.abs(g_full).max()),
                   n_obs=float(buf[-2]), n_lmk=float(buf[-1]), ranges=ranges, O=win.n_obs, L=win.n_lmks))
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,name", [(2, "small"), (3, "C2")])
def test_sharded_reduced_system_adds_up(world, name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, free_port() if r < 0 else PORT, name, q)) for r in range(world)] if False else None
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["err_S"] < 1e-12 and res["err_g"] < 1e-12
    # the shards are a disjoint cover, contiguous and balanced by observation count
    assert res["n_obs"] == res["O"] and res["n_lmk"] == res["L"]
    rg = res["ranges"]
    assert rg[0][0] == 0 and rg[-1][1] == res["L"] and rg[0][2] == 0 and rg[-1][3] == res["O"]
    for a, b in zip(rg[:-1], rg[1:]):
        assert a[1] == b[0] and a[3] == b[2]
    counts = [r[3] - r[2] for r in rg]
    assert max(counts) - min(counts) <= 64


def test_shard_range_edge_cases():
    from sadvio_b200 import synth

    win = synth.make_window("tiny")
    # more ranks than landmarks would leave empty shards: still a cover
    world = 64
    cover = []
    for r in range(world):
        l0, l1, o0, o1 = shard_range(win, r, world)
        assert 0 <= l0 <= l1 <= win.n_lmks and o1 - o0 == int(np.sum((win.obs_lmk >= l0) & (win.obs_lmk < l1)))
        cover.append((l0, l1))
    assert cover[0][0] == 0 and cover[-1][1] == win.n_lmks
    assert all(a[1] == b[0] for a, b in zip(cover[:-1], cover[1:]))
    assert shard_range(win, 0, 1) == (0, win.n_lmks, 0, win.n_obs)

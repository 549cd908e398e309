"""Multi-GPU parity check (run with torchrun on N GPUs): every rank solves the same window with its landmark shard,
the reduced system is all-reduced through NCCL inside the library; rank 0 compares with the single-rank oracle."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sadvio_b200 import api, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    names = sys.argv[1:] or ["small", "C2", "C3"]
    s = api.Solver(device=local)
    api.comm_setup(s, dist, dev)
    for name in names:
        win = synth.make_c4() if name == "C4" else synth.make_window(name)
        dist.barrier()                          # (rank 0 runs the oracle between windows: start every window together)
        rc, d, st = s.solve_window(win)
        s.upload(win)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            s.solve_resident()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        if rank == 0:
            from oracle import oracle
            rc0, d0, st0 = oracle.solve_window(win, nthreads=8)
            rel = lambda a, b: float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))
            print(f"{name} x{world}: iters {st['iterations']} vs {st0['iterations']} term {st['termination']} cost {st['final_cost']:.6f} vs {st0['final_cost']:.6f} "
                  f"pose {rel(d.dpose, d0.dpose):.2e} v {rel(d.dv, d0.dv) if win.vio else 0:.2e} lmk {rel(d.dlmk, d0.dlmk):.2e} "
                  f"solve {dt*1e3:.3f} ms ({st['iterations']/dt:.0f} it/s) cuda graph builds {s.graph_builds()}", flush=True)
            assert st["iterations"] == st0["iterations"] and rel(d.dpose, d0.dpose) < 1e-6 and rel(d.dlmk, d0.dlmk) < 1e-5
    import faulthandler
    faulthandler.dump_traceback_later(20, exit=True)     # (a hang in the teardown shows where)
    t0 = time.perf_counter()
    s.close()
    print(f"rank {rank}: solver closed after {time.perf_counter() - t0:.2f} s", flush=True)
    dist.destroy_process_group()
    print(f"rank {rank}: process group destroyed after {time.perf_counter() - t0:.2f} s", flush=True)
    faulthandler.cancel_dump_traceback_later()


if __name__ == "__main__":
    main()

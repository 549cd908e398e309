"""Generates tests/golden/*.npz: the oracle's solution of seeded synthetic windows (committed regression pins).

    python tests/golden/make_golden.py

The reference itself cannot run in this image (Ceres / Eigen absent, oracle/README.md), so these vectors pin the ORACLE
against drift — its own pin against the reference's known-answer tests is tests/test_oracle_kats.py / ref_fixtures.py.
The CUDA path is compared with them in tests/test_gpu_parity.py::test_golden_solutions (1e-6 relative on the states)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from sadvio_b200 import synth  # noqa: E402

CASES = {"tiny_vio_angular": ("tiny", dict(factor_kind=0)), "tiny_vio_pixel": ("tiny", dict(factor_kind=1)),
         "small_vio_angular": ("small", dict(factor_kind=0))}


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for name, (cfg, kw) in CASES.items():
        win = synth.make_window(cfg, **kw)
        rc, d, st = oracle.solve_window(win, mode=0, nthreads=1)
        assert rc == 0
        np.savez_compressed(os.path.join(here, name + ".npz"), dpose=d.dpose, dv=d.dv, dba=d.dba, dbg=d.dbg, dlmk=d.dlmk,
                            iterations=st["iterations"], termination=st["termination"],
                            trace_cost=np.asarray(st["trace_cost"]), trace_accepted=np.asarray(st["trace_accepted"]),
                            obs_lmk=win.obs_lmk, obs_frame=win.obs_frame, obs_cam=win.obs_cam)
        print(name, "iterations", st["iterations"], "final cost", st["final_cost"])
    # AOptimizer::VIInit on the reference's own initialisation fixture (imu_test.cpp:813-880, tests/ref_fixtures.py: 10 EuRoC keyframes
    # shrunk by 0.5): the oracle's parameter blocks and LM trace
    from tests import ref_fixtures as rf

    win, _ = rf.euroc_viinit_window()
    rc, res, st = oracle.viinit(win, True)
    assert rc == 0
    np.savez_compressed(os.path.join(here, "viinit_euroc.npz"), dv=res["dv"], r_wi=res["r_wi"], lam=res["lam"], scale=res["scale"],
                        iterations=st["iterations"], termination=st["termination"], trace_cost=np.asarray(st["trace_cost"]),
                        trace_accepted=np.asarray(st["trace_accepted"]))
    print("viinit_euroc iterations", st["iterations"], "scale", res["scale"])


if __name__ == "__main__":
    main()

"""Developer probe for k_chol_band: parity of full solves against the oracle and phase cycle counts."""
import sys, os
import numpy as np
sys.path.insert(0, ".")
from sadvio_b200 import synth, api
from oracle import oracle

def rel(a, b):
    return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))

for name in (sys.argv[1:] or ["tiny", "small", "C2", "C3"]):
    w = synth.make_c4() if name == "C4" else synth.make_window(name)
    s = api.Solver()
    rc, d, st = s.solve_window(w)
    rc0, d0, st0 = oracle.solve_window(w, nthreads=8)
    n, npad = s.debug_dims()
    print(f"{name}: n={n} rc {rc}/{rc0} iters {st['iterations']}/{st0['iterations']} term {st['termination']}/{st0['termination']} "
          f"cost {st['final_cost']:.6f}/{st0['final_cost']:.6f} pose {rel(d.dpose, d0.dpose):.2e} lmk {rel(d.dlmk, d0.dlmk):.2e} "
          f"dev {st['ms_solve_device']:.3f} ms", flush=True)
    s.upload(w)
    print(f"  cholesky {s.time_kernel(2, 20)*1e3:.1f} us  schur {s.time_kernel(1, 20)*1e3:.1f} us  lin {s.time_kernel(0, 20)*1e3:.1f} us  backsub {s.time_kernel(3, 20)*1e3:.1f} us")
    prof = s.debug_read(5, 128)
    print("  schur block 0 thread 0 [pre acc gsum inv diag pairs flush] kcycles:", " ".join(f"{x/1e3:7.1f}" for x in prof[64:71]))
    print("  warp 0 [setup wait chain bwd_wait tail backward] kcycles:", " ".join(f"{x/1e3:8.1f}" for x in prof[0:6]))
    print("  chain warp waits [step rhs copy (1,1) (2,1)] kcycles:", " ".join(f"{prof[(q >> 1) * 8 + 6 + (q & 1)]/1e3:7.1f}" for q in range(5)))
    print(f"  hand-over: CTA 0 waited {prof[23]/1e3:.1f} kcycles for CTA 1, separator transfer {prof[31]/1e3:.1f} kcycles")
    print("  backward solve, warp 0 [stage wait, compute, publish, refill] kcycles:", " ".join(f"{prof[(4 + (q >> 1)) * 8 + 6 + (q & 1)]/1e3:7.1f}" for q in range(4)))
    print("  backward solve, CTA 1 warp 0 [stage wait, compute, publish, refill] kcycles:", " ".join(f"{prof[(6 + (q >> 1)) * 8 + 6 + (q & 1)]/1e3:7.1f}" for q in range(4)),
          f" waited {prof[70]/1e3:.1f} for the separator unknowns; CTA 0 waited {prof[71]/1e3:.1f} for CTA 1 at the end")
    pw = prof.reshape(16, 8)
    print("  per warp wait kcycles:", " ".join(f"{x/1e3:6.1f}" for x in pw[:, 1]))
    print("  per warp work kcycles:", " ".join(f"{x/1e3:6.1f}" for x in pw[:, 2]))
    s.close()

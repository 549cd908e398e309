"""Developer tool: time / profile the Jacobian kernel alone on the C5-sized window (run under gpurun, optionally under ncu)."""
import sys
sys.path.insert(0, ".")
from sadvio_b200 import synth, api

name = sys.argv[1] if len(sys.argv) > 1 else "C5"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
w = synth.make_window(name)
s = api.Solver()
s.upload(w)
for k, nm in ((0, "lin_visual"), (1, "lin_schur"), (3, "backsub_cost"), (2, "chol_band")):
    ms = s.time_kernel(k, reps)
    print(f"{name} {nm}: {ms*1e3:.1f} us", end="")
    if k == 0:
        nb = 196 * w.n_obs
        print(f"  {nb/ms/1e6:.0f} GB/s algorithmic", end="")
    print()

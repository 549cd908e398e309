"""Developer tool: one full window solve through the C ABI (run under ncu with SDV_NO_GRAPH=1 for a per-kernel launch list)."""
import sys
sys.path.insert(0, ".")
from sadvio_b200 import synth, api
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
w = synth.make_c4() if name == "C4" else synth.make_window(name)
s = api.Solver()
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    rc, d, st = s.solve_window(w)
print(name, "rc", rc, "iterations", st["iterations"], "device ms", st["ms_solve_device"], "launches", st["kernel_launches"])

"""Developer tool: device time of the resident window solve over many repeats (min / median / mean of the CUDA-event times)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sadvio_b200 import synth, api
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 100
w = synth.make_window(name)
s = api.Solver()
s.upload(w)
t = []
for _ in range(rep + 10):
    rc, st = s.solve_resident()
    t.append(st["ms_solve_device"])
t = np.array(t[10:])
print(f"{name}: {st['iterations']} iterations, device ms min {t.min():.4f} median {np.median(t):.4f} mean {t.mean():.4f}  ({os.environ.get('TAG', '')})")

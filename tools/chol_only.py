"""Developer tool: run a few factor+solve launches on an uploaded window (for ncu captures of the Cholesky kernel)."""
import sys
sys.path.insert(0, ".")
from sadvio_b200 import synth, api
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
w = synth.make_window(name)
s = api.Solver()
s.upload(w)
print(name, "cholesky:", s.time_kernel(2, int(sys.argv[2]) if len(sys.argv) > 2 else 5) * 1e3, "us")

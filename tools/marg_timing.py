"""Developer probe: wall / device / host time of sdv_marginalize on a handle that served a larger window before (capacity-based scratch)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sadvio_b200 import api, synth
win = synth.make_window("C3")
s = api.Solver()
if len(sys.argv) > 1:
    big = synth.make_window(sys.argv[1])
    s.upload(big)
    s.solve_resident()
os.environ["SDV_TIMING"] = "1"
for i in range(3):
    t0 = time.perf_counter()
    dense, sparse, info = s.marginalize(win)
    print(f"call {i}: wall {1e3*(time.perf_counter()-t0):.1f} ms device {info['ms_device']:.2f} host {info['ms_total_host']:.2f} graph builds {s.graph_builds()}", flush=True)

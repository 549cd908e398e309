import sys, time
sys.path.insert(0, ".")
import ctypes as C
from sadvio_b200 import synth, api, abi
w = synth.make_window("C3")
s = api.Solver()
for i in range(6):
    t0=time.perf_counter(); rc,d,st = s.solve_window(w); t1=time.perf_counter()
    print(f"py total {1e3*(t1-t0):.3f} ms; lib host total {st['ms_total_host']:.3f} h2d {st['ms_h2d']:.3f} dev {st['ms_solve_device']:.3f}")
ws = w.as_struct(); out = abi.Delta.zeros(w.n_frames, w.n_lmks); ds = out.as_struct(); stt = abi.SdvStats()
L = api.lib()
for i in range(4):
    t0=time.perf_counter(); L.sdv_solve_window(s._h, C.byref(ws), C.byref(ds), C.byref(stt)); t1=time.perf_counter()
    print(f"raw C call {1e3*(t1-t0):.3f} ms")

"""Developer probe for k_lin_schur (library built with SDV_SCHUR_PROF=1): phase trace of the first tile of CTA 0 and of the last CTA."""
import sys
import numpy as np
sys.path.insert(0, ".")
from sadvio_b200 import synth, api

for name in (sys.argv[1:] or ["C3"]):
    w = synth.make_window(name)
    s = api.Solver()
    s.upload(w)
    print(f"{name}: schur {s.time_kernel(1, 20)*1e3:.1f} us (cold {s.time_kernel(11, 10)*1e3:.1f})  backsub {s.time_kernel(3, 20)*1e3:.1f} us")
    p = s.debug_read(6, 32)
    names = ["tile prologue", "slot operands", "phase A done", "sync", "phase L done", "sync", "phase S done", "sync", "phase B done", "sync", "exit"]
    for off, who in ((0, "CTA 0"), (16, "last CTA")):
        print(f"  {who}: " + "  ".join(f"{n} {p[off + k]/1e3:.1f}k" for k, n in enumerate(names)), f" | globaltimer span {p[off + 12] - p[off + 11]:.0f} ns")
    print(f"  entry skew between CTA 0 and the last CTA: {p[16 + 11] - p[11]:.0f} ns; last exit - first entry: {max(p[12], p[28]) - min(p[11], p[27]):.0f} ns")
    s.close()

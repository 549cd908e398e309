"""One sdv_marginalize of a synthetic window (ncu target): python tools/marg_once.py C3 [repeats]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sadvio_b200 import api, synth

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 1
win = synth.make_window(name)
s = api.Solver()
for _ in range(rep):
    dense, sparse, info = s.marginalize(win, sparsify=True)
    print(f"{name}: m {info['m']} n {info['n']} n_full {info['n_full']} sweeps {info['eig_sweeps_m']}/{info['eig_sweeps_n']} device {info['ms_device']:.3f} ms host {info['ms_total_host']:.3f} ms")

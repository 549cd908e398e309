"""One sdv_viinit on the keyframes / IMU pairs of a synthetic window (ncu target): python tools/viinit_once.py C3"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sadvio_b200 import api, synth
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
win = synth.make_window(name)
s = api.Solver()
for _ in range(2):
    rc, res, st = s.viinit(win, True)
print(name, "rc", rc, "iterations", st["iterations"], "n", st["n_reduced"], "device ms", st["ms_solve_device"])

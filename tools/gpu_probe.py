"""Developer probe: compare the CUDA path with the oracle stage by stage (run under gpurun)."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from sadvio_b200 import synth, abi, api
from oracle import oracle

def rel(a, b):
    return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))

def probe(name, **kw):
    print(f"==== {name} {kw}", flush=True)
    w = synth.make_window(name, **kw)
    s = api.Solver()
    s.upload(w)
    n, npad = s.debug_dims()
    print("dims", n, npad, "F", w.n_frames, "L", w.n_lmks, "O", w.n_obs)
    # --- visual factors at x = 0 and at a random x
    rng = np.random.default_rng(0)
    for tag, x in (("x=0", None), ("x=rand", abi.Delta(rng.normal(0, 0.02, (w.n_frames, 6)), rng.normal(0, 0.02, (w.n_frames, 3)),
                                                     rng.normal(0, 0.01, (w.n_frames, 3)), rng.normal(0, 0.001, (w.n_frames, 3)),
                                                     rng.normal(0, 0.05, (w.n_lmks, 3))))):
        if x is not None:
            nf = w.n_fixed
            if nf:
                x.dpose[-nf:] = 0; x.dv[-nf:] = 0; x.dba[-nf:] = 0; x.dbg[-nf:] = 0
        r, Jp, Jl, c = s.eval_visual(x)
        r0, Jp0, Jl0, c0 = oracle.eval_visual(w, x)
        print(f" visual {tag}: r {rel(r, r0):.2e} Jp {rel(Jp, Jp0):.2e} Jl {rel(Jl, Jl0):.2e} cost {c:.6f} vs {c0:.6f}")
        if w.vio:
            ri, Ji, rb = s.eval_imu(x)
            ri0, Ji0, rb0 = oracle.eval_imu(w, x)
            print(f" imu    {tag}: r {rel(ri, ri0):.2e} J {rel(Ji, Ji0):.2e} rbias {rel(rb, rb0):.2e}")
    # --- full solve
    t0 = time.time()
    rc, st = s.solve_resident()
    d = s.download()
    t1 = time.time()
    rc0, d0, st0 = oracle.solve_window(w, mode=0, nthreads=8)
    t2 = time.time()
    print(" gpu :", rc, st["iterations"], st["termination"], st["initial_cost"], st["final_cost"], "fixed", st["fixed_cost"])
    print(" orc :", rc0, st0["iterations"], st0["termination"], st0["initial_cost"], st0["final_cost"], "fixed", st0["fixed_cost"])
    print(" gpu trace", [f"{c:.6g}" for c in st["trace_cost"]], st["trace_accepted"])
    print(" orc trace", [f"{c:.6g}" for c in st0["trace_cost"]], st0["trace_accepted"])
    print(" gpu radius", [f"{c:.4g}" for c in st["trace_radius"]])
    print(" orc radius", [f"{c:.4g}" for c in st0["trace_radius"]])
    print(" gpu model", [f"{c:.6g}" for c in st["trace_model_change"]])
    print(" orc model", [f"{c:.6g}" for c in st0["trace_model_change"]])
    print(f" delta: pose {rel(d.dpose, d0.dpose):.2e} v {rel(d.dv, d0.dv):.2e} ba {rel(d.dba, d0.dba):.2e} bg {rel(d.dbg, d0.dbg):.2e} lmk {rel(d.dlmk, d0.dlmk):.2e}")
    print(f" time gpu solve {st['ms_solve_device']:.3f} ms device, {st['ms_total_host']:.3f} ms host, launches {st['kernel_launches']}; oracle {1e3*(t2-t1):.1f} ms")
    for k, nm in ((0, "lin_visual"), (1, "schur"), (2, "cholesky")):
        print(f"  kernel {nm}: {s.time_kernel(k, 20)*1e3:.1f} us")
    prof = s.debug_read(5, 128).reshape(16, 8)
    print("  chol phase cycles per rank [wait_Lkk load_Lkk trsm lookahead|publish wait_col|ld+diag update|chol backward]:")
    for r in (0, 1, 7, 15):
        print("   rank", r, " ".join(f"{x/1e3:8.1f}k" for x in prof[r, :7]))
    import ctypes as C
    mic = np.zeros(72)
    api.lib().sdv_debug_micro.argtypes = [C.c_void_p, abi.c_double_p]
    api.lib().sdv_debug_micro(s._h, mic.ctypes.data_as(abi.c_double_p))
    for i, nm in enumerate(["chol32_reg", "chol32_hyb", "trsm32_reg", "trsm32_smem", "diag_update", "dmma_tile", "load_row32", "store_tile"]):
        print(f"  micro {nm:12s} cycles/rep:", " ".join(f"{x:7.0f}" for x in mic[i * 8:i * 8 + 5]))
    print("  update tile cycles, 1/2/4/8 warps per CTA: dmma", " ".join(f"{x:7.0f}" for x in mic[64:68]), "| dfma", " ".join(f"{x:7.0f}" for x in mic[68:72]))
    s.close()

if __name__ == "__main__":
    names = sys.argv[1:] or ["tiny", "small", "C2"]
    for nm in names:
        if ":" in nm:
            a, b = nm.split(":")
            probe(a, factor_kind=int(b))
        else:
            probe(nm)

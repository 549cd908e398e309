"""Build the sm_100a shared library in-tree (sadvio_b200/_lib/libsadvio_b200.so).

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "_lib")
# SDV_LIB selects a prebuilt experiment library (tools/gpu_round.sh variant mode); it is never rebuilt from here.
LIB = os.environ.get("SDV_LIB") or os.path.join(LIB_DIR, "libsadvio_b200.so")
SOURCES = ["sdv_lib.cu"]
HEADERS = ["sdv_kernels.cuh", "sdv_fused.cuh", "sdv_chol.cuh", "sdv_chol_band.cuh", "sdv_math.cuh", "sdv_types.cuh", "sdv_preint.cuh", "sdv_marg.cuh", "sdv_marg_host.cuh", "sdv_viinit.cuh", "sdv_peer.cuh", "sdv_struct.cuh", os.path.join("..", "..", "include", "sdv.h")]
NVCC_FLAGS = [
    *(["-DSDV_BAND_PROF"] if os.environ.get("SDV_BAND_PROF") else []),
    *(["-DSDV_SCHUR_PROF"] if os.environ.get("SDV_SCHUR_PROF") else []),
    # k_chol_band switches (sdv_chol_band.cuh): SDV_BAND_BABE=0 / SDV_BAND_BACKWARD_V2=0 build the round-1 kernel,
    # SDV_BAND_REV=1 (with SDV_BAND_BABE=0) the index-reversed debugging variant, SDV_BAND_STREAM1=0 the tensor-core first-column updates
    *([f"-D{k}={os.environ[k]}" for k in ("SDV_BAND_BACKWARD_V2", "SDV_BAND_BACKWARD_V3", "SDV_BAND_REV", "SDV_BAND_BABE", "SDV_BAND_STREAM1", "SDV_BAND_SKEW", "SDV_FT", "SDV_FT_LMK", "SDV_FUSED_CPS") if os.environ.get(k)]),
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared",
]


def is_stale() -> bool:
    if os.environ.get("SDV_LIB"):
        return False
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES], "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))

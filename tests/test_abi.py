"""Host-side checks that need no GPU: the C-ABI library loads, exports every symbol include/sdv.h declares, the ctypes
mirror matches the C struct layout, and the product path refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from sadvio_b200 import abi, api, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sdv.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdv_[a-z_0-9]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    lib_path = build.build()
    lib = C.CDLL(lib_path)
    names = declared_functions()
    assert {"sdv_create", "sdv_solve_window", "sdv_destroy", "sdv_strerror", "sdv_upload_window", "sdv_solve_resident",
            "sdv_download_delta", "sdv_eval_visual", "sdv_eval_imu", "sdv_comm_init"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sdv.h but not exported"
    assert lib.sdv_abi_version() == abi.SDV_ABI_VERSION


def test_ctypes_mirror_matches_c_layout():
    probe = r"""
    #include <stdio.h>
    #include <stddef.h>
    #include "sdv.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu\n", sizeof(sdv_config), sizeof(sdv_dense_prior), sizeof(sdv_sparse_prior),
             sizeof(sdv_window), sizeof(sdv_delta), sizeof(sdv_stats));
      printf("%zu %zu %zu %zu\n", offsetof(sdv_window, T_f_w), offsetof(sdv_window, obs_lmk), offsetof(sdv_window, imu_cov),
             offsetof(sdv_window, sparse_prior));
      printf("%zu %zu %zu\n", offsetof(sdv_stats, trace_accepted), offsetof(sdv_stats, kernel_launches),
             offsetof(sdv_sparse_prior, l2l_sqrt_inf));
      printf("%zu %zu %zu %zu\n", sizeof(sdv_imu_intervals), sizeof(sdv_preint), offsetof(sdv_imu_intervals, eta), offsetof(sdv_imu_intervals, rate_hz));
      printf("%zu %zu %zu %zu %zu\n", sizeof(sdv_viinit_result), offsetof(sdv_viinit_result, r_wi), offsetof(sdv_viinit_result, lambda),
             offsetof(sdv_viinit_result, R_w_i), offsetof(sdv_viinit_result, scale));
      printf("%zu %zu %zu %zu\n", sizeof(sdv_marginal_sizes), sizeof(sdv_marginal), offsetof(sdv_marginal, lmk_sqrt_inf), offsetof(sdv_marginal, l2l_delta));
      return 0; }
    """
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "p.c")
        open(src, "w").write(probe)
        exe = os.path.join(td, "p")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    got = [int(x) for x in out]
    want = [C.sizeof(abi.SdvConfig), C.sizeof(abi.SdvDensePrior), C.sizeof(abi.SdvSparsePrior), C.sizeof(abi.SdvWindow),
            C.sizeof(abi.SdvDelta), C.sizeof(abi.SdvStats),
            abi.SdvWindow.T_f_w.offset, abi.SdvWindow.obs_lmk.offset, abi.SdvWindow.imu_cov.offset, abi.SdvWindow.sparse_prior.offset,
            abi.SdvStats.trace_accepted.offset, abi.SdvStats.kernel_launches.offset, abi.SdvSparsePrior.l2l_sqrt_inf.offset,
            C.sizeof(abi.SdvImuIntervals), C.sizeof(abi.SdvPreint), abi.SdvImuIntervals.eta.offset, abi.SdvImuIntervals.rate_hz.offset,
            C.sizeof(abi.SdvViinitResult), abi.SdvViinitResult.r_wi.offset, abi.SdvViinitResult.lambda_.offset, abi.SdvViinitResult.R_w_i.offset,
            abi.SdvViinitResult.scale.offset,
            C.sizeof(abi.SdvMarginalSizes), C.sizeof(abi.SdvMarginal), abi.SdvMarginal.lmk_sqrt_inf.offset, abi.SdvMarginal.l2l_delta.offset]
    assert got == want


def test_default_config_is_the_reference_options():
    cfg = api.default_config()
    assert cfg.max_num_iterations == 20 and cfg.function_tolerance == 1e-3      # AOptimizer.cpp:380,384
    assert cfg.initial_trust_region_radius == 1e4 and cfg.jacobi_scaling == 1   # Ceres 2.2 defaults
    assert cfg.min_lm_diagonal == 1e-6 and cfg.max_lm_diagonal == 1e32 and cfg.min_relative_decrease == 1e-3


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly; on a GPU box this test is a no-op."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.BackendUnavailable):
        api.Solver()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sadvio_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt.replace("oracle/README", ""), f

"""CPU: the C-ABI library loads, exports every symbol include/brr_gpu.h declares, and refuses to
compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import bayesrrcmd_b200.capi as capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from bayesrrcmd_b200 import build
    build.build_cuda()
    return capi.load_library()


def test_header_symbols_are_exported(lib):
    header = open(os.path.join(ROOT, "include", "brr_gpu.h")).read()
    declared = set(re.findall(r"\b(brr_gpu_[a-z0-9_]+)\s*\(", header))
    declared -= {"brr_gpu_ctx", "brr_gpu_config", "brr_gpu_status", "brr_gpu_layout", "brr_gpu_sweep_out"}
    assert declared == set(capi.SYMBOLS)
    for name in sorted(declared):
        assert hasattr(lib, name), name


def test_abi_version(lib):
    assert lib.brr_gpu_abi_version() == capi.ABI_VERSION


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.BrrGpuError, match="no CPU fallback"):
        capi.Engine(100, 10, 4)


def test_bad_config_is_rejected(lib):
    cfg = capi.Config(abi_version=999, n_inds=10, n_markers=1, n_components=4, n_groups=1, world=1)
    h = ctypes.c_void_p()
    assert lib.brr_gpu_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"ABI version" in lib.brr_gpu_last_error()


def test_product_never_imports_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "bayesrrcmd_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "ref_lib" not in txt and "bayesr_oracle" not in txt, f

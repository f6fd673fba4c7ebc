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


def test_sharded_sparse_layouts_are_rejected(lib):
    """The ragged / eigen sweeps have no cross-shard exchange: world > 1 (or a partial shard) must fail loudly at create,
    before any device is touched, instead of sampling from partial dots."""
    for layout in (capi.LAYOUT_RAGGED, capi.LAYOUT_EIGEN):
        for kw in (dict(world=2, rank=0, n_inds=100, n_inds_total=200), dict(world=1, rank=0, n_inds=100, n_inds_total=200)):
            cfg = capi.Config(abi_version=capi.ABI_VERSION, layout=layout, n_markers=4, n_components=4, n_groups=1, **kw)
            h = ctypes.c_void_p()
            assert lib.brr_gpu_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
            assert b"no cross-shard exchange" in lib.brr_gpu_last_error()
    cfg = capi.Config(abi_version=capi.ABI_VERSION, layout=0, n_markers=4, n_components=4, n_groups=1, world=2, rank=0,
                      n_inds=3000000, n_inds_total=6000000)
    h = ctypes.c_void_p()
    assert lib.brr_gpu_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"accumulators" in lib.brr_gpu_last_error()


def test_numpy_twin_of_the_synthetic_generator(lib):
    """oracle/synth_np.py (what bench.py's reference arm uses, so that it maps no product library) reproduces the product's
    host twin of the device generator bit for bit."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import synth_np
    for N, first, cnt, lu, pm in [(1003, 0, 20, False, 0.01), (4096, 5, 7, True, 0.001), (64, 100000, 3, False, 0.0)]:
        a = capi.synth_bed_host(42, N, first, cnt, 0.01, 0.5, lu, pm)
        b = synth_np.synth_bed(42, N, np.arange(first, first + cnt), 0.01, 0.5, lu, pm)
        assert a.shape == b.shape and (a == b).all()


def test_product_never_imports_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "bayesrrcmd_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "ref_lib" not in txt and "bayesr_oracle" not in txt, f

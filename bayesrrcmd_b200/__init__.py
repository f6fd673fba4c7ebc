"""bayesrrcmd_b200: B200-native (sm_100a) engine for the BayesR Gibbs-sweep hot path of
ctggroup/BayesRRcmd, behind the C ABI declared in include/brr_gpu.h.

Python here is plumbing only (ctypes binding, torch.distributed launch glue for bench/tests); the
product is csrc/ (CUDA kernels + C ABI) and host/ (C++ mirror of the reference's
Analysis / AnalysisGraph / Kernel / Marker / MarkerBuilder / MarkerCache interfaces).
"""
from .capi import Engine, BrrGpuError, load_library, LIB_PATH  # noqa: F401

"""TEST INFRASTRUCTURE ONLY: ctypes access to oracle/liboracle.so (bayesr_oracle.c, the plain-C
restatement of the reference's BayesR hot path).  Importable only from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class Model(C.Structure):
    _fields_ = [("N", C.c_uint32), ("K", C.c_uint32), ("nGroups", C.c_uint32), ("cva", C.c_void_p),
                ("pi", C.c_void_p), ("sigmaG", C.c_void_p), ("sigmaE", C.c_double)]


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])


def load():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        lib.orc_dense_compute_num.restype = C.c_double
        lib.orc_ragged_compute_num.restype = C.c_double
        lib.orc_eigen_compute_num.restype = C.c_double
        lib.orc_init_epsilon.restype = C.c_double
        lib.orc_mu_mean_dense.restype = C.c_double
        lib.orc_dense_compute_num.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_double]
        lib.orc_dense_eps_update.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, C.c_double]
        lib.orc_ragged_compute_num.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                               C.c_void_p, C.c_uint32, C.c_double, C.c_double]
        lib.orc_ragged_eps_update.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                              C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_double,
                                              C.c_double, C.c_void_p]
        lib.orc_eigen_compute_num.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, C.c_double]
        lib.orc_eigen_eps_update.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double,
                                             C.c_double, C.c_void_p]
        lib.orc_sample_marker.argtypes = [C.POINTER(Model), C.c_int, C.c_double, C.c_double, C.c_double,
                                          C.c_double, C.c_void_p, C.c_void_p]
        lib.orc_sweep_dense.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                                        C.POINTER(Model), C.c_void_p, C.c_uint32] + [C.c_void_p] * 6
        lib.orc_sweep_ragged.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                                         C.POINTER(Model), C.c_void_p, C.c_uint32] + [C.c_void_p] * 7
        lib.orc_sweep_eigen.argtypes = lib.orc_sweep_ragged.argtypes
        lib.orc_init_epsilon.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        lib.orc_mu_mean_dense.argtypes = [C.c_void_p, C.c_uint32, C.c_double]
        lib.orc_mu_apply.argtypes = [C.c_void_p, C.c_uint32, C.c_double, C.c_double]
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def decode(bedcol, N):
    cls = np.empty(N, np.uint8)
    load().orc_decode_bed_column(_p(bedcol), C.c_uint32(N), _p(cls))
    return cls


def dense_column(bedcol, N):
    cx = np.empty(N)
    load().orc_dense_column(_p(bedcol), C.c_uint32(N), _p(cx))
    return cx


def dense_lut(bedcol, N):
    lut = np.empty(4)
    load().orc_dense_lut(_p(bedcol), C.c_uint32(N), _p(lut))
    return lut


def ragged_marker(bedcol, N):
    stats = np.empty(4)
    ones = np.empty(N, np.int32); twos = np.empty(N, np.int32); miss = np.empty(N, np.int32)
    n = [C.c_uint32() for _ in range(3)]
    load().orc_ragged_marker(_p(bedcol), C.c_uint32(N), _p(stats), _p(ones), C.byref(n[0]), _p(twos),
                             C.byref(n[1]), _p(miss), C.byref(n[2]))
    return stats, ones[:n[0].value].copy(), twos[:n[1].value].copy(), miss[:n[2].value].copy()


def eigen_stats(bedcol, N):
    stats = np.empty(4)
    load().orc_eigen_stats(_p(bedcol), C.c_uint32(N), _p(stats))
    return stats


def dense_compute_num(cx, eps, beta_old):
    return load().orc_dense_compute_num(_p(cx), _p(eps), len(cx), beta_old)


def dense_eps_update(cx, eps, beta_old, beta):
    load().orc_dense_eps_update(_p(cx), _p(eps), len(cx), beta_old, beta)


def ragged_compute_num(stats, ones, twos, eps, beta_old, eps_sum):
    return load().orc_ragged_compute_num(_p(stats), _p(ones), len(ones), _p(twos), len(twos), _p(eps),
                                         len(eps), beta_old, eps_sum)


def ragged_eps_update(stats, ones, twos, miss, eps, beta_old, beta, eps_sum):
    es = C.c_double(eps_sum)
    load().orc_ragged_eps_update(_p(stats), _p(ones), len(ones), _p(twos), len(twos), _p(miss), len(miss),
                                 _p(eps), len(eps), beta_old, beta, C.byref(es))
    return es.value


def eigen_compute_num(stats, cls, eps, beta_old, eps_sum):
    return load().orc_eigen_compute_num(_p(stats), _p(cls), _p(eps), len(eps), beta_old, eps_sum)


def eigen_eps_update(stats, cls, eps, beta_old, beta, eps_sum):
    es = C.c_double(eps_sum)
    load().orc_eigen_eps_update(_p(stats), _p(cls), _p(eps), len(eps), beta_old, beta, C.byref(es))
    return es.value


class ModelArrays:
    """Keeps the numpy buffers an orc_model points to alive."""

    def __init__(self, N, cva, pi, sigmaG, sigmaE):
        self.cva = np.ascontiguousarray(np.atleast_2d(cva), dtype=np.float64)
        self.pi = np.ascontiguousarray(np.atleast_2d(pi), dtype=np.float64)
        self.sigmaG = np.ascontiguousarray(np.atleast_1d(sigmaG), dtype=np.float64)
        nG, km1 = self.cva.shape
        assert self.pi.shape == (nG, km1 + 1) and self.sigmaG.shape == (nG,)
        self.c = Model(N=N, K=km1 + 1, nGroups=nG, cva=self.cva.ctypes.data, pi=self.pi.ctypes.data,
                       sigmaG=self.sigmaG.ctypes.data, sigmaE=float(sigmaE))


def sample_marker(model, group, num, beta_old, u, z):
    b = C.c_double(); a = C.c_double()
    k = load().orc_sample_marker(C.byref(model.c), group, num, beta_old, u, z, C.byref(b), C.byref(a))
    return k, b.value, a.value


def sweep(layout, bed, M, order, rand_uz, model, groups, eps, beta, comp, acum, tokens=0, eps_sum=0.0):
    """In-place sweep on float64 arrays; returns (v, betasq, eps_sum)."""
    order = np.ascontiguousarray(order, dtype=np.uint32)
    rand_uz = np.ascontiguousarray(rand_uz, dtype=np.float64)
    g = None if groups is None else np.ascontiguousarray(groups, dtype=np.int32)
    v = np.zeros((model.c.nGroups, model.c.K)); bs = np.zeros(model.c.nGroups)
    bed = np.ascontiguousarray(bed, dtype=np.uint8)
    if layout == "dense":
        rc = load().orc_sweep_dense(_p(bed), M, _p(order), len(order), _p(rand_uz), C.byref(model.c), _p(g),
                                    tokens, _p(eps), _p(beta), _p(comp), _p(acum), _p(v), _p(bs))
    else:
        es = C.c_double(eps_sum)
        fn = load().orc_sweep_eigen if layout == "eigen" else load().orc_sweep_ragged
        rc = fn(_p(bed), M, _p(order), len(order), _p(rand_uz), C.byref(model.c), _p(g),
                                     tokens, _p(eps), C.byref(es), _p(beta), _p(comp), _p(acum), _p(v), _p(bs))
        eps_sum = es.value
    assert rc == 0, rc
    return v, bs, eps_sum


def init_epsilon(y):
    y = np.ascontiguousarray(y, dtype=np.float32)
    eps = np.empty(len(y))
    sE = load().orc_init_epsilon(_p(y), len(y), _p(eps))
    return eps, sE

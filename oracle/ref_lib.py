"""TEST INFRASTRUCTURE ONLY: ctypes access to oracle/_ref/libbrr_ref*.so (the unmodified reference
compiled by oracle/Makefile) and readers for the reference's fixture formats.

Only tests/, oracle/make_golden.py and bench.py's `--impl reference` / cpu_baseline leg import this.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DP = C.POINTER(C.c_double)


class ChainDump(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "mu", "sigmaE_in", "sigmaG_in", "pi_in", "rand_uz", "order", "eps_in", "beta_in", "epssum_in",
        "beta_out", "comp_out", "acum_out", "eps_out", "v_out", "betasq_out", "epssum_out",
        "sigmaE_post", "sigmaG_post", "pi_post", "gamma_post", "eps_post")]


class ChainArgs(C.Structure):
    _fields_ = [("bed", C.c_void_p), ("N", C.c_uint32), ("M", C.c_uint32), ("y", C.c_void_p),
                ("groups", C.c_void_p), ("nGroups", C.c_uint32), ("S", C.c_void_p), ("K", C.c_uint32),
                ("seed", C.c_uint32), ("iterations", C.c_uint32), ("layout", C.c_int32),
                ("dump", C.c_void_p), ("execSeconds", C.c_double), ("totalSeconds", C.c_double),
                ("buildSeconds", C.c_double), ("graphKind", C.c_int32), ("analysisTokens", C.c_uint32),
                ("X", C.c_void_p), ("nF", C.c_uint32)]


def ref_available(fast=False):
    return os.path.exists(_path(fast))


def _path(fast):
    if fast == "gpu":       # the strict reference build + the GpuSweepGraph of oracle/gpu_graph_ref.cpp, linked against the product library
        return os.path.join(HERE, "_ref", "libbrr_ref_gpu.so")
    return os.path.join(HERE, "_ref", "libbrr_ref_fast.so" if fast else "libbrr_ref.so")


_libs = {}


def load(fast=False):
    if fast not in _libs:
        lib = C.CDLL(_path(fast))
        lib.ref_dense_column.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        lib.ref_ragged_marker.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 7
        lib.ref_eigen_stats.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        if hasattr(lib, "ref_write_pp_files"):     # (a library built before the function existed still serves the other tests)
            lib.ref_write_pp_files.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_char_p, C.c_char_p]
        lib.ref_marker_ops.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_double,
                                       C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_run_chain.argtypes = [C.POINTER(ChainArgs)]
        _libs[fast] = lib
    return _libs[fast]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def dense_column(bedcol, N):
    out = np.empty(N)
    assert load().ref_dense_column(_p(bedcol), N, _p(out)) == 0
    return out


def ragged_marker(bedcol, N):
    stats = np.empty(4)
    ones = np.empty(N, np.int32); twos = np.empty(N, np.int32); miss = np.empty(N, np.int32)
    n = [C.c_uint32() for _ in range(3)]
    assert load().ref_ragged_marker(_p(bedcol), N, _p(stats), _p(ones), C.byref(n[0]), _p(twos),
                                    C.byref(n[1]), _p(miss), C.byref(n[2])) == 0
    return stats, ones[:n[0].value].copy(), twos[:n[1].value].copy(), miss[:n[2].value].copy()


def eigen_stats(bedcol, N):
    stats = np.empty(4)
    assert load().ref_eigen_stats(_p(bedcol), N, _p(stats)) == 0
    return stats


def write_pp_files(bed, N, M, layout, compress, marker_file, index_file):
    """The reference's preprocessed marker file + index file for a PLINK payload (its own Marker::write / compress)."""
    bed = np.ascontiguousarray(bed, dtype=np.uint8)
    rc = load().ref_write_pp_files(_p(bed), bed.shape[1], N, M, layout, int(compress), marker_file.encode(), index_file.encode())
    if rc:
        raise RuntimeError("ref_write_pp_files failed: %d" % rc)


def marker_ops(layout, bedcol, N, eps, beta_old, beta_new, eps_sum=0.0):
    """-> (num, delta_eps or None, epsilonSum field after calculateEpsilonChange)."""
    num = C.c_double(); es = C.c_double(eps_sum)
    delta = np.full(N, np.nan)
    eps = np.ascontiguousarray(eps, dtype=np.float64)
    rc = load().ref_marker_ops(layout, _p(bedcol), N, _p(eps), beta_old, beta_new, C.byref(num),
                               _p(delta), C.byref(es))
    assert rc == 0, rc
    return num.value, delta, es.value


class _Quiet:
    """The reference prints per-iteration progress to stdout; keep it out of callers' stdout."""

    def __enter__(self):
        import sys
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        C.CDLL(None).fflush(None)
        os.dup2(self.saved, 1)
        os.close(self.null); os.close(self.saved)


def run_chain(bed, N, M, y, S, seed, iterations, layout=0, groups=None, dump=True, fast=False, graph="sequential", tokens=20,
              keep=None, X=None):
    """Run the reference's BayesRBase::runGibbs; returns (dict of per-iteration arrays, timings).
    graph="parallel": through the reference's own ParallelGraph (asyncppbayes: TBB, all cores, `tokens` markers in flight).
    keep: names of the per-iteration arrays to record (None = all); the others share one scratch row (long chains)."""
    S = np.ascontiguousarray(np.atleast_2d(S), dtype=np.float64)
    nG, K = S.shape[0], S.shape[1] + 1
    bed = np.ascontiguousarray(bed, dtype=np.uint8)
    y = np.ascontiguousarray(y, dtype=np.float32)
    g = None if groups is None else np.ascontiguousarray(groups, dtype=np.int32)
    it = iterations
    shapes = dict(mu=(it,), sigmaE_in=(it,), sigmaG_in=(it, nG), pi_in=(it, nG, K), rand_uz=(it, M, 2),
                  order=(it, M), eps_in=(it, N), beta_in=(it, M), epssum_in=(it,), beta_out=(it, M),
                  comp_out=(it, M), acum_out=(it, M), eps_out=(it, N), v_out=(it, nG, K),
                  betasq_out=(it, nG), epssum_out=(it,), sigmaE_post=(it,), sigmaG_post=(it, nG),
                  pi_post=(it, nG, K))
    Xf = None
    if X is not None:                                        # fixed-effect covariates (N x nF), float32 column-major as Data::X
        Xf = np.asfortranarray(np.atleast_2d(X), dtype=np.float32)
        shapes["gamma_post"] = (it, Xf.shape[1]); shapes["eps_post"] = (it, N)
    arrays = {}
    d = ChainDump()
    if dump:
        assert keep is None, "keep= needs the harness to index per iteration: record everything (dump=True) on short chains instead"
        for k, shp in shapes.items():
            arrays[k] = np.zeros(shp, dtype=np.uint32 if k == "order" else np.float64)
            setattr(d, k, arrays[k].ctypes.data)
    a = ChainArgs(bed=bed.ctypes.data, N=N, M=M, y=y.ctypes.data,
                  groups=None if g is None else g.ctypes.data, nGroups=nG, S=S.ctypes.data, K=K,
                  seed=seed, iterations=iterations, layout=layout,
                  dump=C.addressof(d) if dump else None, analysisTokens=tokens,
                  X=None if Xf is None else Xf.ctypes.data, nF=0 if Xf is None else Xf.shape[1],
                  graphKind={"sequential": 0, "parallel": 1, "gpu": 2, "gpu-async": 3}[graph])
    if graph.startswith("gpu"):
        fast = "gpu"
    C.CDLL(None).srand(1)   # glibc rand() drives std::random_shuffle; fresh-process state
    lib = load(fast)
    with _Quiet():
        rc = lib.ref_run_chain(C.byref(a))
    assert rc == 0, rc
    return arrays, dict(exec_s=a.execSeconds, total_s=a.totalSeconds, build_s=a.buildSeconds)


# ---- fixture readers (semantics of src/data.cpp:85-133, 334-365, 424-453) ---------------------

def read_fam(path):
    ids = []
    with open(path) as f:
        for line in f:
            t = line.split()
            if len(t) >= 2:
                ids.append(t[0] + ":" + t[1])
    return ids


def read_bim_count(path):
    with open(path) as f:
        return sum(1 for line in f if line.strip())


def read_bed(path, N, M):
    raw = np.fromfile(path, dtype=np.uint8)
    assert raw[0] == 0x6c and raw[1] == 0x1b and raw[2] == 0x01, "not a SNP-major PLINK .bed"
    cb = (N + 3) // 4
    return raw[3:3 + cb * M].reshape(M, cb).copy()


def read_phen(path, fam_ids):
    """y filled in PHENOTYPE-FILE order for lines whose FID:IID is known and value != NA, stored as
    float32 (src/data.cpp:334-365, src/data.hpp:92)."""
    known = set(fam_ids)
    y = np.zeros(len(fam_ids), dtype=np.float32)
    n = 0
    with open(path) as f:
        for line in f:
            t = line.split()
            if len(t) < 3:
                continue
            if (t[0] + ":" + t[1]) in known and t[2] != "NA":
                try:
                    v = float(t[2])
                except ValueError:
                    v = 0.0          # atof semantics
                y[n] = np.float32(v)
                n += 1
    return y


def read_groups(path):
    g = []
    with open(path) as f:
        for line in f:
            t = line.split()
            if len(t) >= 2:
                g.append(int(t[1]))
    return np.asarray(g, dtype=np.int32)

"""ctypes binding of include/brr_gpu.h.  Fails loudly when the CUDA library is missing: there is
no CPU fallback and nothing here ever touches oracle/."""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BRR_GPU_LIB") or os.path.join(PKG, "libbrr_gpu.so")   # BRR_GPU_LIB: another build of the same ABI (A/B timing)
ABI_VERSION = 1
LAYOUT_DENSE2BIT, LAYOUT_RAGGED, LAYOUT_EIGEN = 0, 1, 2
_LAYOUTS = {"dense": 0, "ragged": 1, "eigen": 2}
IPC_HANDLE_BYTES = 64

# every symbol include/brr_gpu.h declares (tests check the library exports exactly these)
SYMBOLS = [
    "brr_gpu_last_error", "brr_gpu_abi_version", "brr_gpu_create", "brr_gpu_destroy",
    "brr_gpu_upload_dense2bit", "brr_gpu_get_counts", "brr_gpu_finalize_markers", "brr_gpu_set_dense_lut",
    "brr_gpu_get_dense_lut", "brr_gpu_upload_ragged", "brr_gpu_upload_ragged_from_bed", "brr_gpu_get_ragged_stats",
    "brr_gpu_set_model", "brr_gpu_set_state", "brr_gpu_get_state", "brr_gpu_eps_add_scalar", "brr_gpu_eps_reduce",
    "brr_gpu_sweep", "brr_gpu_sweep_stage", "brr_gpu_sweep_run", "brr_gpu_sweep_fetch", "brr_gpu_compute_num",
    "brr_gpu_apply_delta", "brr_gpu_ipc_export", "brr_gpu_ipc_attach", "brr_gpu_attach_local",
    "brr_gpu_num_blocks", "brr_gpu_stream", "brr_gpu_synth_dense", "brr_gpu_synth_ragged", "brr_gpu_synth_bed_host", "brr_gpu_info",
    "brr_gpu_launch_count", "brr_gpu_profile_cycles", "brr_gpu_kernel_version", "brr_gpu_profile_cycles5", "brr_gpu_profile_cycles6", "brr_gpu_profile_ctas6", "brr_gpu_last_sweep_info",
]


class BrrGpuError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("device", C.c_int32), ("layout", C.c_int32), ("n_inds", C.c_uint32),
                ("n_inds_total", C.c_uint32), ("n_markers", C.c_uint32), ("n_components", C.c_uint32),
                ("n_groups", C.c_uint32), ("block_markers", C.c_uint32), ("max_ctas", C.c_uint32),
                ("rank", C.c_int32), ("world", C.c_int32)]


class SweepOut(C.Structure):
    _fields_ = [("eps_sum", C.c_double), ("eps_sqnorm", C.c_double), ("n_nonskipped", C.c_uint32),
                ("n_blocks", C.c_uint32), ("kernel_ms", C.c_float), ("n_launches", C.c_uint32)]


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BrrGpuError("%s is missing: build it with `python -m bayesrrcmd_b200.build` "
                          "(CUDA extension required, no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, u32, dbl = C.c_void_p, C.c_uint32, C.c_double
    lib.brr_gpu_last_error.restype = C.c_char_p
    lib.brr_gpu_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.brr_gpu_destroy.argtypes = [vp]
    lib.brr_gpu_destroy.restype = None
    lib.brr_gpu_upload_dense2bit.argtypes = [vp, vp, C.c_size_t, u32, u32]
    lib.brr_gpu_get_counts.argtypes = [vp, vp]
    lib.brr_gpu_finalize_markers.argtypes = [vp, vp]
    lib.brr_gpu_set_dense_lut.argtypes = [vp, vp]
    lib.brr_gpu_get_dense_lut.argtypes = [vp, vp]
    lib.brr_gpu_upload_ragged.argtypes = [vp, vp, vp, vp]
    lib.brr_gpu_upload_ragged_from_bed.argtypes = [vp, vp, C.c_size_t]
    lib.brr_gpu_get_ragged_stats.argtypes = [vp, vp]
    lib.brr_gpu_set_model.argtypes = [vp, vp, vp]
    lib.brr_gpu_set_state.argtypes = [vp, vp, vp, vp, vp]
    lib.brr_gpu_get_state.argtypes = [vp, vp, vp, vp, vp]
    lib.brr_gpu_eps_add_scalar.argtypes = [vp, dbl, C.POINTER(dbl)]
    lib.brr_gpu_eps_reduce.argtypes = [vp, C.POINTER(dbl), C.POINTER(dbl)]
    lib.brr_gpu_sweep.argtypes = [vp, vp, u32, vp, vp, vp, dbl, dbl, u32, vp, vp, C.POINTER(SweepOut)]
    lib.brr_gpu_sweep_stage.argtypes = [vp, vp, u32, vp, vp, vp, dbl, dbl, u32]
    lib.brr_gpu_sweep_run.argtypes = [vp]
    lib.brr_gpu_sweep_fetch.argtypes = [vp, vp, vp, C.POINTER(SweepOut)]
    lib.brr_gpu_compute_num.argtypes = [vp, u32, dbl, dbl, C.POINTER(dbl)]
    lib.brr_gpu_apply_delta.argtypes = [vp, u32, dbl, dbl, C.POINTER(dbl)]
    lib.brr_gpu_ipc_export.argtypes = [vp, vp]
    lib.brr_gpu_ipc_attach.argtypes = [vp, vp]
    lib.brr_gpu_attach_local.argtypes = [C.POINTER(vp), C.c_int]
    lib.brr_gpu_num_blocks.argtypes = [vp]
    lib.brr_gpu_num_blocks.restype = u32
    lib.brr_gpu_stream.argtypes = [vp]
    lib.brr_gpu_stream.restype = vp
    lib.brr_gpu_synth_dense.argtypes = [vp, C.c_uint64, dbl, dbl, C.c_int, dbl, u32]
    lib.brr_gpu_synth_ragged.argtypes = [vp, C.c_uint64, dbl, dbl, C.c_int, dbl, u32]
    lib.brr_gpu_synth_bed_host.argtypes = [C.c_uint64, dbl, dbl, C.c_int, dbl, u32, u32, u32, u32, vp]
    lib.brr_gpu_info.argtypes = [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(C.c_uint64)]
    lib.brr_gpu_launch_count.argtypes = [vp]
    lib.brr_gpu_launch_count.restype = C.c_uint64
    lib.brr_gpu_profile_cycles.argtypes = [vp, vp, C.c_int]
    lib.brr_gpu_kernel_version.restype = C.c_char_p
    lib.brr_gpu_profile_cycles5.argtypes = [vp, vp, C.c_int]
    lib.brr_gpu_profile_cycles6.argtypes = [vp, vp, C.c_int]
    lib.brr_gpu_profile_ctas6.argtypes = [vp, vp, C.c_uint32, vp, C.c_int]
    lib.brr_gpu_last_sweep_info.argtypes = [vp, C.POINTER(u32), C.POINTER(u32)]
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class Engine:
    """Thin object wrapper over a brr_gpu_ctx.  All arrays are host numpy arrays."""

    def __init__(self, n_inds, n_markers, n_components, n_groups=1, layout="dense", device=0, block_markers=0,
                 max_ctas=0, n_inds_total=None, rank=0, world=1):
        self.lib = load_library()
        self.N, self.M, self.K, self.nG = int(n_inds), int(n_markers), int(n_components), int(n_groups)
        self.layout = layout
        cfg = Config(abi_version=ABI_VERSION, device=device,
                     layout=_LAYOUTS[layout], n_inds=self.N,
                     n_inds_total=self.N if n_inds_total is None else int(n_inds_total), n_markers=self.M,
                     n_components=self.K, n_groups=self.nG, block_markers=block_markers, max_ctas=max_ctas,
                     rank=rank, world=world)
        h = C.c_void_p()
        self._h = None
        self._check(self.lib.brr_gpu_create(C.byref(cfg), C.byref(h)))
        self._h = h

    def _check(self, rc):
        if rc != 0:
            raise BrrGpuError("brr_gpu error %d: %s" % (rc, self.lib.brr_gpu_last_error().decode()))

    def close(self):
        if self._h is not None:
            self.lib.brr_gpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # column supply
    def upload_bed(self, bed, first=0):
        bed = np.ascontiguousarray(bed, dtype=np.uint8)
        assert bed.ndim == 2
        if self.layout == "dense":
            self._check(self.lib.brr_gpu_upload_dense2bit(self._h, _ptr(bed), bed.shape[1], first, bed.shape[0]))
        else:
            assert first == 0 and bed.shape[0] == self.M
            self._check(self.lib.brr_gpu_upload_ragged_from_bed(self._h, _ptr(bed), bed.shape[1]))

    def upload_ragged(self, offsets, idx, stats=None):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        self._check(self.lib.brr_gpu_upload_ragged(self._h, _ptr(offsets), _ptr(idx), _ptr(_f64(stats))))

    def counts(self):
        c = np.empty((self.M, 4), dtype=np.uint32)
        self._check(self.lib.brr_gpu_get_counts(self._h, _ptr(c)))
        return c

    def finalize(self, global_counts=None):
        g = None if global_counts is None else np.ascontiguousarray(global_counts, dtype=np.uint32)
        self._check(self.lib.brr_gpu_finalize_markers(self._h, _ptr(g)))

    def set_dense_lut(self, lut):
        self._check(self.lib.brr_gpu_set_dense_lut(self._h, _ptr(_f64(lut))))

    def dense_lut(self):
        lut = np.empty((self.M, 4))
        self._check(self.lib.brr_gpu_get_dense_lut(self._h, _ptr(lut)))
        return lut

    def ragged_stats(self):
        s = np.empty((self.M, 4))
        self._check(self.lib.brr_gpu_get_ragged_stats(self._h, _ptr(s)))
        return s

    def synth_dense(self, seed, maf_lo=0.01, maf_hi=0.5, log_uniform=False, p_missing=0.001, ind_offset=0):
        self._check(self.lib.brr_gpu_synth_dense(self._h, seed, maf_lo, maf_hi, int(log_uniform), p_missing, ind_offset))

    def synth_ragged(self, seed, maf_lo=0.01, maf_hi=0.5, log_uniform=False, p_missing=0.001, ind_offset=0):
        self._check(self.lib.brr_gpu_synth_ragged(self._h, seed, maf_lo, maf_hi, int(log_uniform), p_missing, ind_offset))

    # model / state
    def set_model(self, cva, groups=None):
        cva = _f64(np.atleast_2d(cva))
        assert cva.shape == (self.nG, self.K - 1)
        g = None if groups is None else np.ascontiguousarray(groups, dtype=np.int32)
        self._check(self.lib.brr_gpu_set_model(self._h, _ptr(g), _ptr(cva)))

    def set_state(self, eps=None, beta=None, comp=None, acum=None):
        a = [_f64(x) for x in (eps, beta, comp, acum)]
        self._check(self.lib.brr_gpu_set_state(self._h, *[_ptr(x) for x in a]))

    def get_state(self, eps=True, beta=True, comp=True, acum=True, out=None):
        """out: optional dict of caller-owned float64 arrays (e.g. pinned host memory) for "eps" / "beta" / "comp" / "acum"."""
        out = out or {}
        def buf(name, want, n):
            if not want:
                return None
            x = out.get(name)
            if x is None:
                return np.empty(n)
            assert x.dtype == np.float64 and x.flags.c_contiguous and x.size == n
            return x
        e = buf("eps", eps, self.N); b = buf("beta", beta, self.M); c = buf("comp", comp, self.M); a = buf("acum", acum, self.M)
        self._check(self.lib.brr_gpu_get_state(self._h, _ptr(e), _ptr(b), _ptr(c), _ptr(a)))
        return e, b, c, a

    def eps_add_scalar(self, c):
        s = C.c_double()
        self._check(self.lib.brr_gpu_eps_add_scalar(self._h, c, C.byref(s)))
        return s.value

    def eps_reduce(self):
        s = C.c_double(); q = C.c_double()
        self._check(self.lib.brr_gpu_eps_reduce(self._h, C.byref(s), C.byref(q)))
        return s.value, q.value

    # sweep
    def sweep(self, order, rand_uz, pi, sigmaG, sigmaE, eps_sum=0.0, async_tokens=0):
        order, rand_uz, pi, sigmaG = self._sweep_args(order, rand_uz, pi, sigmaG)
        v = np.empty((self.nG, self.K)); bs = np.empty(self.nG); out = SweepOut()
        self._check(self.lib.brr_gpu_sweep(self._h, _ptr(order), len(order), _ptr(rand_uz), _ptr(pi), _ptr(sigmaG),
                                           sigmaE, eps_sum, async_tokens, _ptr(v), _ptr(bs), C.byref(out)))
        return v, bs, out

    def _sweep_args(self, order, rand_uz, pi, sigmaG):
        order = np.ascontiguousarray(order, dtype=np.uint32)
        rand_uz = _f64(rand_uz); pi = _f64(np.atleast_2d(pi)); sigmaG = _f64(np.atleast_1d(sigmaG))
        assert rand_uz.shape == (self.M, 2) and pi.shape == (self.nG, self.K) and sigmaG.shape == (self.nG,)
        return order, rand_uz, pi, sigmaG

    def sweep_stage(self, order, rand_uz, pi, sigmaG, sigmaE, eps_sum=0.0, async_tokens=0):
        order, rand_uz, pi, sigmaG = self._sweep_args(order, rand_uz, pi, sigmaG)
        self._check(self.lib.brr_gpu_sweep_stage(self._h, _ptr(order), len(order), _ptr(rand_uz), _ptr(pi),
                                                 _ptr(sigmaG), sigmaE, eps_sum, async_tokens))

    def sweep_run(self):
        self._check(self.lib.brr_gpu_sweep_run(self._h))

    def sweep_fetch(self):
        v = np.empty((self.nG, self.K)); bs = np.empty(self.nG); out = SweepOut()
        self._check(self.lib.brr_gpu_sweep_fetch(self._h, _ptr(v), _ptr(bs), C.byref(out)))
        return v, bs, out

    # parity hooks
    def compute_num(self, marker, beta_old, eps_sum=0.0):
        n = C.c_double()
        self._check(self.lib.brr_gpu_compute_num(self._h, marker, beta_old, eps_sum, C.byref(n)))
        return n.value

    def apply_delta(self, marker, beta_old, beta_new):
        d = C.c_double()
        self._check(self.lib.brr_gpu_apply_delta(self._h, marker, beta_old, beta_new, C.byref(d)))
        return d.value

    # multi-GPU
    def ipc_export(self):
        buf = (C.c_ubyte * IPC_HANDLE_BYTES)()
        self._check(self.lib.brr_gpu_ipc_export(self._h, buf))
        return bytes(buf)

    def ipc_attach(self, handles):
        blob = b"".join(handles)
        self._check(self.lib.brr_gpu_ipc_attach(self._h, blob))

    @staticmethod
    def attach_local(engines):
        """Wire several engines of this process (rank i = engines[i]) for the in-kernel exchange."""
        arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
        engines[0]._check(engines[0].lib.brr_gpu_attach_local(arr, len(engines)))

    def info(self):
        a, b, c, d = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint64()
        self._check(self.lib.brr_gpu_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(n_ctas=a.value, slice_inds=b.value, block_markers=c.value, genotype_bytes=d.value,
                    kernel_version=self.lib.brr_gpu_kernel_version().decode())

    def launch_count(self):
        return int(self.lib.brr_gpu_launch_count(self._h))

    PROFILE_PHASES = ["wait", "xchg", "readback", "sample", "join", "publish", "lazy_cross", "lazy_mma", "lazy_barrier",
                      "lazy_readback", "solve", "resample", "finalize", "carry", "axpy", "quantize", "tail", "prologue",
                      "walk_iters", "lazy_rounds", "lazy_rows", "blocks_updated", "worker_wait", "worker_dot",
                      "pub_reduce", "pub_missing", "pub_arrive", "lz_loop", "lz_red", "pub_pre", "lz_c0", "lz_c1", "lz_c2",
                      "w_dot", "w_missing", "w_combine", "spare3"]
    PROFILE_COUNTERS = ("walk_iters", "lazy_rounds", "lazy_rows", "blocks_updated")
    PROFILE_OVERLAPPED = ("worker_wait", "worker_dot", "w_dot", "w_missing", "w_combine")     # timed on a dot-pass warp, concurrent with the walker phases

    def profile_cycles(self, reset=True):
        """Phase cycle counters of the dense sweep kernel (needs BRR_GPU_PROFILE=1 at construction)."""
        a = np.zeros(40, dtype=np.uint64)
        self._check(self.lib.brr_gpu_profile_cycles(self._h, _ptr(a), int(reset)))
        return dict(zip(self.PROFILE_PHASES, (int(x) for x in a)))

    PROFILE5 = ["w_plan", "w_bar", "w_read", "w_sample", "w_fix", "w_file", "d_gate", "d_apply", "d_tma", "d_loop", "d_post", "d_comb",
                "blocks", "gram_blocks", "restarts", "applies", "p0_plan", "p0_issue"]

    def profile_cycles5(self, reset=True):
        """Phase cycle counters of the version-5 dense kernel (walker warp 0 / first dot warp / planner warp of CTA 0)."""
        a = np.zeros(24, dtype=np.uint64)
        self._check(self.lib.brr_gpu_profile_cycles5(self._h, _ptr(a), int(reset)))
        return dict(zip(self.PROFILE5, (int(x) for x in a)))

    PROFILE6 = ["d_dec", "d_apply", "d_full", "d_loop", "d_miss", "d_cross", "d_acc", "pu_wait", "pu_work", "pr_plan", "pr_empty", "pr_issue",
                "c_poll", "c_plan", "c_decide", "w_cnt", "w_read", "w_pred", "w_sample", "w_fix", "w_valid", "blocks", "walks", "applies", "restarts",
                "gram_blocks", "row_blocks", "d_pre", "d_fullonly", "d_loopgram", "c_iters"]

    def profile_cycles6(self, reset=True):
        """Phase cycle counters of the version-6 dense kernel (one warp per role of CTA 0)."""
        a = np.zeros(32, dtype=np.uint64)
        self._check(self.lib.brr_gpu_profile_cycles6(self._h, _ptr(a), int(reset)))
        return dict(zip(self.PROFILE6, (int(x) for x in a)))

    PROFILE6_CTA = ["dec_wait", "update", "rows_wait", "mma_loop", "acc_wait", "store", "blocks", "smid"]

    def profile_ctas6(self, reset=True):
        """Version-6 dense kernel: the first dot warp's cycles per CTA -> uint64 array [nCTA, 8] (columns: PROFILE6_CTA)."""
        a = np.zeros((1024, 8), dtype=np.uint64); n = C.c_uint32()
        self._check(self.lib.brr_gpu_profile_ctas6(self._h, _ptr(a), 1024, C.byref(n), int(reset)))
        return a[:n.value]

    def last_sweep_info(self):
        """-> (dense kernel version of the last sweep: 4 or 5, pipeline restarts of a version-5 sweep)."""
        k, r = C.c_uint32(), C.c_uint32()
        self._check(self.lib.brr_gpu_last_sweep_info(self._h, C.byref(k), C.byref(r)))
        return k.value, r.value

    def stream(self):
        return self.lib.brr_gpu_stream(self._h)


def synth_bed_host(seed, n_inds, first, count, maf_lo=0.01, maf_hi=0.5, log_uniform=False, p_missing=0.001, ind_offset=0):
    lib = load_library()
    bed = np.zeros((count, (n_inds + 3) // 4), dtype=np.uint8)
    rc = lib.brr_gpu_synth_bed_host(seed, maf_lo, maf_hi, int(log_uniform), p_missing, n_inds, ind_offset, first, count, _ptr(bed))
    if rc != 0:
        raise BrrGpuError(lib.brr_gpu_last_error().decode())
    return bed

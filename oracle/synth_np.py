"""TEST INFRASTRUCTURE ONLY: numpy twin of the bench's counter-based synthetic genotypes.

The engine generates its synthetic store on the device (brr_gpu_synth_dense; generator in
bayesrrcmd_b200/csrc/store_kernels.cuh: mix64 / synth_class, thresholds in brr_gpu.cu:synth_thresholds).
The reference arm of bench.py must see the SAME data set without mapping any product library into
its process, so the generator is restated here on uint64 numpy arrays; tests/test_synth_twin.py pins
it bit for bit against the product's host twin (brr_gpu_synth_bed_host).
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(x):
    x = x.copy()
    x ^= x >> np.uint64(30); x *= np.uint64(0xbf58476d1ce4e5b9)
    x ^= x >> np.uint64(27); x *= np.uint64(0x94d049bb133111eb)
    x ^= x >> np.uint64(31)
    return x


def thresholds(seed, markers, maf_lo, maf_hi, log_uniform, p_missing):
    """-> (t0, t1, tmiss) uint32 arrays, one entry per marker id in `markers`."""
    with np.errstate(over="ignore"):
        mk = np.asarray(markers, dtype=np.uint64)
        h = _mix64(np.uint64(seed) * np.uint64(0x2545F4914F6CDD1D) + mk * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0x1234567))
    u = (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    maf = maf_lo * np.power(maf_hi / maf_lo, u) if log_uniform else maf_lo + (maf_hi - maf_lo) * u
    p0 = (1.0 - maf) * (1.0 - maf); p1 = 2.0 * maf * (1.0 - maf)
    two32 = 4294967296.0
    t0 = np.minimum(4294967295.0, np.floor(p0 * two32)).astype(np.uint32)
    t1 = np.minimum(4294967295.0, np.floor((p0 + p1) * two32)).astype(np.uint32)
    tm = np.full(len(mk), min(4294967295.0, np.floor(p_missing * two32)), dtype=np.float64).astype(np.uint32)
    return t0, t1, tm


def classes(seed, marker, n_inds, t0, t1, tm, ind_offset=0):
    """Allele-count classes (0, 1, 2; 3 = missing) of one marker for individuals ind_offset .. ind_offset+n_inds-1."""
    with np.errstate(over="ignore"):
        ind = np.arange(ind_offset, ind_offset + n_inds, dtype=np.uint64)
        h = _mix64((np.uint64(seed) ^ (np.uint64(marker) * np.uint64(0x9E3779B97F4A7C15))) + ind * np.uint64(0xD1B54A32D192ED03))
    r = (h >> np.uint64(32)).astype(np.uint32); r2 = (h & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    cls = np.where(r < t0, 0, np.where(r < t1, 1, 2)).astype(np.uint8)
    cls[r2 < tm] = 3
    return cls


_PLINK = np.array([3, 2, 0, 1], dtype=np.uint8)     # class 0,1,2,missing -> PLINK 11,10,00,01


def synth_bed(seed, n_inds, markers, maf_lo=0.01, maf_hi=0.5, log_uniform=False, p_missing=0.001, ind_offset=0):
    """PLINK .bed payload (no magic), one column of ceil(n_inds/4) bytes per entry of `markers` (marker ids)."""
    markers = np.asarray(markers, dtype=np.uint64)
    t0, t1, tm = thresholds(seed, markers, maf_lo, maf_hi, log_uniform, p_missing)
    cb = (n_inds + 3) // 4
    bed = np.zeros((len(markers), cb), dtype=np.uint8)
    pad = cb * 4 - n_inds
    for i, mk in enumerate(markers):
        code = _PLINK[classes(seed, mk, n_inds, t0[i], t1[i], tm[i], ind_offset)]
        if pad:
            code = np.concatenate([code, np.zeros(pad, np.uint8)])
        c4 = code.reshape(cb, 4)
        bed[i] = c4[:, 0] | (c4[:, 1] << 2) | (c4[:, 2] << 4) | (c4[:, 3] << 6)
    return bed

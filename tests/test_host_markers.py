"""Column supply of the host side (bayesrrcmd_b200/host/markers.*): Marker / MarkerBuilder / MarkerCache and the reference's
preprocessed file formats (.ppbed, .ragged.sparsebed, .eigen.sparsebed + index files, plain and zlib-compressed).

CPU: files written by the REFERENCE's own markers (oracle/_ref: Marker::write / Marker::compress, the reference's index writer)
are read back through the host façade and compared, marker by marker, with what the reference holds in memory; the façade's own
builders and writers reproduce those files; a standardised dense column goes back to its 2-bit codes and value table.
GPU: an engine fed from the preprocessed files through MarkerCache::upload runs the same chain as an engine fed from the .bed.
"""
import ctypes as C
import os

import numpy as np
import pytest

import ref_lib as R
from helpers import Golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "bayesrrcmd_b200")
LAYOUTS = {"dense": 0, "ragged": 1, "eigen": 2}
EXT = {0: ".ppbed", 1: ".ragged.sparsebed", 2: ".eigen.sparsebed"}


@pytest.fixture(scope="module")
def host():
    from bayesrrcmd_b200 import build
    build.build_all()
    lib = C.CDLL(os.path.join(PKG, "libbrr_host.so"))
    vp = C.c_void_p
    lib.brr_host_cache_from_files.argtypes = [C.c_char_p, C.c_int, C.c_uint, C.c_uint, C.c_int]
    lib.brr_host_cache_from_bed.argtypes = [vp, C.c_size_t, C.c_int, C.c_uint, C.c_uint]
    lib.brr_host_cache_write.argtypes = [C.c_char_p, C.c_int, C.c_int]
    lib.brr_host_cache_marker.argtypes = [C.c_uint] + [vp] * 11
    lib.brr_host_cache_upload.argtypes = [vp]
    lib.brr_host_dense_column_to_codes.argtypes = [vp, C.c_uint, vp, vp]
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _small(name="missing_small", m=40):
    g = Golden(name)
    M = min(g.M, m)
    return g, np.ascontiguousarray(g.bed[:M]), g.N, M


def _cached_marker(host, i, N, layout):
    dense = np.zeros(N); stats = np.zeros(4)
    lists = [np.zeros(N, np.int32) for _ in range(3)]; counts = [C.c_uint(0) for _ in range(3)]
    values = np.zeros(N); indices = np.zeros(N, np.int32); nnz = C.c_uint(0)
    rc = host.brr_host_cache_marker(i, _p(dense), _p(stats), _p(lists[0]), C.byref(counts[0]), _p(lists[1]), C.byref(counts[1]),
                                    _p(lists[2]), C.byref(counts[2]), _p(values), _p(indices), C.byref(nnz))
    assert rc == 0
    return dict(dense=dense, stats=stats, lists=[l[:c.value] for l, c in zip(lists, counts)], values=values[:nnz.value], indices=indices[:nnz.value])


needs_ref = pytest.mark.skipif(not R.ref_available(), reason="oracle/_ref is not built")


@needs_ref
@pytest.mark.parametrize("layout", ["dense", "ragged", "eigen"])
@pytest.mark.parametrize("compress", [0, 1])
def test_reads_the_references_preprocessed_files(host, tmp_path, layout, compress):
    g, bed, N, M = _small()
    L = LAYOUTS[layout]
    base = str(tmp_path / "data")
    R.write_pp_files(bed, N, M, L, compress, base + EXT[L], base + EXT[L] + "index")
    assert host.brr_host_cache_from_files((base + ".bed").encode(), L, M, N, compress) == 0
    idx = np.fromfile(base + EXT[L] + "index", dtype=np.uint64).reshape(M, 3)
    assert (idx[1:, 0] == np.cumsum(idx[:-1, 1])).all()                # pos = running sum of the stored sizes
    for j in range(M):
        got = _cached_marker(host, j, N, L)
        if layout == "dense":
            assert np.array_equal(got["dense"], R.dense_column(bed[j], N))          # the file's doubles, bit for bit
        elif layout == "ragged":
            stats, ones, twos, missing = R.ragged_marker(bed[j], N)
            assert np.array_equal(got["stats"], stats)
            for a, b in zip(got["lists"], (ones, twos, missing)):
                assert np.array_equal(a, b)
        else:
            assert np.array_equal(got["stats"], R.eigen_stats(bed[j], N))
            stats, ones, twos, missing = R.ragged_marker(bed[j], N)
            nz = np.sort(np.r_[ones, twos, missing])
            assert np.array_equal(got["indices"], nz)
            val = np.where(np.isin(nz, ones), 1.0, np.where(np.isin(nz, twos), 2.0, got["stats"][0]))
            assert np.array_equal(got["values"], val)


@needs_ref
@pytest.mark.parametrize("layout", ["dense", "ragged", "eigen"])
def test_own_builders_and_writers_reproduce_the_files(host, tmp_path, layout):
    g, bed, N, M = _small()
    L = LAYOUTS[layout]
    ref_base, own_base = str(tmp_path / "ref"), str(tmp_path / "own")
    R.write_pp_files(bed, N, M, L, 0, ref_base + EXT[L], ref_base + EXT[L] + "index")
    assert host.brr_host_cache_from_bed(_p(bed), bed.shape[1], L, M, N) == 0
    assert host.brr_host_cache_write((own_base + ".bed").encode(), L, 0) == 0
    ref_idx = np.fromfile(ref_base + EXT[L] + "index", dtype=np.uint64); own_idx = np.fromfile(own_base + EXT[L] + "index", dtype=np.uint64)
    assert np.array_equal(ref_idx, own_idx)
    if layout == "dense":                                               # the column mean is summed in another order: last-bit differences
        a = np.fromfile(ref_base + EXT[L]); b = np.fromfile(own_base + EXT[L])
        assert np.max(np.abs(a - b)) <= 1e-13
    else:
        assert open(ref_base + EXT[L], "rb").read() == open(own_base + EXT[L], "rb").read()
    # compressed: the same content after inflation (read back through the cache)
    assert host.brr_host_cache_write((own_base + "z.bed").encode(), L, 1) == 0
    assert host.brr_host_cache_from_files((own_base + "z.bed").encode(), L, M, N, 1) == 0
    j = M // 2
    got = _cached_marker(host, j, N, L)
    if layout == "dense":
        assert np.max(np.abs(got["dense"] - R.dense_column(bed[j], N))) <= 1e-13
    else:
        assert np.array_equal(got["stats"], R.ragged_marker(bed[j], N)[0] if layout == "ragged" else R.eigen_stats(bed[j], N))


@needs_ref
def test_dense_column_back_to_codes(host):
    g, bed, N, M = _small("missing_small", 60)
    for j in range(M):
        col = R.dense_column(bed[j], N)
        out = np.zeros((N + 3) // 4, np.uint8); lut = np.zeros(4)
        assert host.brr_host_dense_column_to_codes(_p(col), N, _p(out), _p(lut)) == 0
        codes = (out[np.arange(N) // 4] >> (2 * (np.arange(N) % 4))) & 3
        cls = np.array([2, 3, 1, 0])[codes]                             # PLINK bit pair -> allele count (3 = missing)
        assert np.array_equal(lut[cls], col)                            # the table reproduces the column exactly
        orig = np.array([2, 3, 1, 0])[(bed[j][np.arange(N) // 4] >> (2 * (np.arange(N) % 4))) & 3]
        same = cls == orig
        assert same.all() or np.allclose(lut[cls[~same]], lut[orig[~same]], atol=1e-9)   # (a class whose value coincides with the imputed one)


def _pack(codes):
    """(M, N) PLINK bit pairs (0 = two copies, 1 = missing, 2 = one copy, 3 = none) -> .bed payload rows."""
    M, N = codes.shape
    out = np.zeros((M, (N + 3) // 4), np.uint8)
    for i in range(N):
        out[:, i // 4] |= codes[:, i].astype(np.uint8) << (2 * (i % 4))
    return out


@needs_ref
@pytest.mark.parametrize("N", [2, 3, 5, 17, 65])
def test_ragged_sizes_and_sparse_columns_against_the_reference(host, tmp_path, N):
    """Individual counts that are not multiples of 4, columns with one informative genotype among missing ones, all-heterozygous columns:
    the files of the host's builders equal the reference's, and the reference's files are read back, for every layout, plain and deflated."""
    rng = np.random.default_rng(N)
    cols = [np.r_[0, np.full(N - 1, 1)], np.r_[2, np.full(N - 1, 1)], np.r_[0, np.full(N - 1, 3)], np.r_[2, 0, np.full(N - 2, 3)][:N],
            np.r_[0, rng.integers(0, 4, N - 1)], np.r_[2, rng.choice([0, 1, 3], N - 1)], np.full(N, 2)]
    bed = _pack(np.array(cols)); M = len(cols)
    for name, L in LAYOUTS.items():
        for compress in (0, 1):
            ref, own = str(tmp_path / ("ref_%s%d" % (name, compress))), str(tmp_path / ("own_%s%d" % (name, compress)))
            R.write_pp_files(bed, N, M, L, compress, ref + EXT[L], ref + EXT[L] + "index")
            assert host.brr_host_cache_from_bed(_p(bed), bed.shape[1], L, M, N) == 0
            assert host.brr_host_cache_write((own + ".bed").encode(), L, compress) == 0
            a, b = open(ref + EXT[L], "rb").read(), open(own + EXT[L], "rb").read()
            if name == "dense" and not compress:
                assert np.allclose(np.frombuffer(a), np.frombuffer(b), atol=1e-13, equal_nan=True)
            elif name != "dense":
                assert a == b and open(ref + EXT[L] + "index", "rb").read() == open(own + EXT[L] + "index", "rb").read()
            assert host.brr_host_cache_from_files((ref + ".bed").encode(), L, M, N, compress) == 0
            if name == "ragged":
                for j in range(M):
                    got = _cached_marker(host, j, N, L)
                    stats, ones, twos, missing = R.ragged_marker(bed[j], N)
                    assert np.array_equal(got["stats"], stats, equal_nan=True)
                    for x, y in zip(got["lists"], (ones, twos, missing)):
                        assert np.array_equal(x, y)


def test_a_marker_that_does_not_vary_is_refused_like_the_reference_does(host, tmp_path):
    """src/raggedsparsemarker.cpp:68-77, src/eigensparsemarker.cpp:59-71: a sparse marker without a single 1 or 2 is not valid -- reading
    a file that holds one fails (the reference prints "SNPs that do not vary ... should be removed prior to analysis")."""
    N = 9
    bed = _pack(np.array([np.r_[0, np.full(N - 1, 3)], np.full(N, 3), np.r_[2, np.full(N - 1, 3)]]))     # the second column: nobody carries the allele
    for L in (1, 2):
        base = str(tmp_path / ("mono%d" % L))
        assert host.brr_host_cache_from_bed(_p(bed), bed.shape[1], L, 3, N) == 0
        assert host.brr_host_cache_write((base + ".bed").encode(), L, 0) == 0
        assert host.brr_host_cache_from_files((base + ".bed").encode(), L, 3, N, 0) != 0


def test_missing_files_are_reported(host, tmp_path):
    assert host.brr_host_cache_from_files(str(tmp_path / "nothing.bed").encode(), 1, 5, 10, 0) != 0


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("layout", ["dense", "ragged", "eigen"])
def test_engine_fed_from_preprocessed_files_runs_the_same_chain(host, tmp_path, layout):
    from bayesrrcmd_b200 import Engine
    g, bed, N, M = _small("missing_small", 10 ** 9)
    L = LAYOUTS[layout]
    base = str(tmp_path / "data")
    R.write_pp_files(bed, N, M, L, 1, base + EXT[L], base + EXT[L] + "index")
    assert host.brr_host_cache_from_files((base + ".bed").encode(), L, M, N, 1) == 0
    rng = np.random.default_rng(4)
    eps0 = rng.normal(size=N); pi = rng.dirichlet(np.r_[20.0, np.ones(g.K - 1)], size=g.nG); sigmaG = rng.uniform(0.2, 0.8, g.nG)
    results = []
    for fed_from_files in (False, True):
        e = Engine(N, M, g.K, g.nG, layout=layout)
        if fed_from_files:
            assert host.brr_host_cache_upload(e._h) == 0
        else:
            e.upload_bed(bed); e.finalize()
        e.set_model(g.S, g.groups)
        e.set_state(eps=eps0, beta=np.zeros(M), comp=np.zeros(M), acum=np.zeros(M))
        r2 = np.random.default_rng(9)
        es = float(eps0.sum())
        for _ in range(3):
            order = r2.permutation(M).astype(np.uint32)
            v, bs, out = e.sweep(order, np.c_[r2.random(M), r2.normal(size=M)], pi, sigmaG, 0.6, es, 0)
            es = out.eps_sum
        results.append(e.get_state())
    for a, b in zip(*results):
        assert np.max(np.abs(a - b)) <= 1e-8
    assert np.array_equal(results[0][2], results[1][2])                 # components

"""The C++ host mirror (bayesrrcmd_b200/host: Options / Data / GpuSweepGraph / GpuBayesR::runGibbs / SampleWriter) through
its `brr`-style command line.

CPU: the binary builds, parses the reference's command line, reads PLINK files, and stops with "no CPU fallback" when there is
no GPU.  GPU: a full chain FROM THE SEED -- the host draws every random number itself (std::mt19937 + std:: distributions, the
mapping the oracle's reference build uses, and std::random_shuffle for the marker order) -- must reproduce the reference chain
recorded in tests/golden (mu, sigmaE, sigmaG, beta, components, acum, epsilon within 1e-8): end-to-end parity of the drop-in.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from helpers import Golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "bayesrrcmd_b200")
EXE = os.path.join(PKG, "brr_gpu")


@pytest.fixture(scope="module")
def built():
    from bayesrrcmd_b200 import build
    build.build_all()
    assert os.path.exists(EXE)
    return EXE


def write_plink(g, d, groups_file=False):
    """PLINK bed/bim/fam + phenotype (+ group) files of a golden fixture."""
    base = os.path.join(d, "data")
    with open(base + ".bed", "wb") as f:
        f.write(bytes([0x6c, 0x1b, 0x01]))
        f.write(np.ascontiguousarray(g.bed).tobytes())
    with open(base + ".fam", "w") as f:
        for i in range(g.N):
            f.write("F%d I%d 0 0 0 -9\n" % (i, i))
    with open(base + ".bim", "w") as f:
        for j in range(g.M):
            f.write("1 snp%d 0 %d A C\n" % (j, j + 1))
    with open(base + ".phen", "w") as f:
        for i in range(g.N):
            f.write("F%d I%d %.9g\n" % (i, i, float(g.y[i])))
    if groups_file:
        with open(base + ".group", "w") as f:
            for j in range(g.M):
                f.write("snp%d %d\n" % (j, int(g.groups[j])))
    return base


def run(exe, args, **kw):
    return subprocess.run([exe] + args, capture_output=True, text=True, timeout=600, **kw)


def test_host_library_exports_entry_point(built):
    lib = ctypes.CDLL(os.path.join(PKG, "libbrr_host.so"))
    assert hasattr(lib, "brr_host_main")


def test_command_line_errors(built, tmp_path):
    r = run(built, ["--no-such-option"])
    assert r.returncode == 1 and "invalid option" in r.stderr
    r = run(built, ["--analysis-type", "gauss", "--data-file", "x.bed", "--pheno", "x.phen"])
    assert r.returncode == 1 and "ppbayes" in r.stderr
    r = run(built, ["--preprocess", "--data-file", str(tmp_path / "missing.bed")])
    assert r.returncode == 1 and "can not open" in r.stderr
    r = run(built, ["--analysis-type", "ppbayes", "--data-file", str(tmp_path / "missing.bed"), "--pheno", "x.phen"])
    assert r.returncode == 1 and "can not open" in r.stderr
    r = run(built, ["--analysis-type", "ppbayes", "--S", "0.1,zzz"])
    assert r.returncode == 1 and "variance component" in r.stderr


def test_options_from_a_file(built, tmp_path):
    """--inp-file (src/options.cpp:280-324): `key value` pairs; a pair whose key starts with # or // is a comment; unknown keys are errors."""
    f = tmp_path / "run.txt"
    f.write_text("#note x\nanalysisType ppbayes\nchainLength 3\nburnin 0\nthin 1\nseed 5\nS 0.1,0.01\nphenotypeFile %s\n" % (tmp_path / "x.phen"))
    r = run(built, ["--inp-file", str(f)])
    assert r.returncode == 1 and "--data-file" in r.stderr            # options accepted; the data file is what is missing
    f.write_text("nonsense 3\n")
    r = run(built, ["--inp-file", str(f)])
    assert r.returncode == 1 and "invalid option nonsense" in r.stderr
    r = run(built, ["--inp-file", str(tmp_path / "none.txt")])
    assert r.returncode == 1 and "can not open" in r.stderr
    for opt in ("--iterLog", "--colLog", "--out"):
        r = run(built, [opt])
        assert r.returncode == 1 and "needs a value" in r.stderr


def test_reads_plink_files_then_requires_a_gpu(built, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    g = Golden("missing_small")
    base = write_plink(g, str(tmp_path), groups_file=True)
    r = run(built, ["--analysis-type", "ppbayes", "--data-file", base + ".bed", "--pheno", base + ".phen", "--group", base + ".group",
                    "--S", "0.01,0.001;0.01,0.001", "--chain-length", "2", "--seed", "7", "--mcmc-samples", str(tmp_path / "o.csv")])
    assert "%d individuals to be included" % g.N in r.stdout
    assert "%d SNPs to be included" % g.M in r.stdout
    assert "Groups read from file: 2" in r.stdout
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


def read_samples(path, M, N, nG, nF=0):
    rows = np.loadtxt(path, delimiter=",", skiprows=1, ndmin=2)
    out = []
    for r in rows:
        o = 2
        d = dict(iteration=int(r[0]), mu=r[1], beta=r[o:o + M]); o += M
        d["sigmaE"] = r[o]; o += 1
        d["sigmaG"] = r[o:o + nG]; o += nG
        d["gamma"] = r[o:o + nF]; o += nF
        d["comp"] = r[o:o + M]; o += M
        d["acum"] = r[o:o + M]; o += M
        d["eps"] = r[o:o + N]
        out.append(d)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name,layout,sargs", [("uk10k", "dense", []), ("uk10k", "ragged", ["--sparse-data", "ragged"])])
def test_chain_from_seed_reproduces_the_reference(built, tmp_path, name, layout, sargs):
    g = Golden(name)
    chain = g.chain(layout)
    base = write_plink(g, str(tmp_path))
    out = str(tmp_path / "samples.csv")
    iters = len(chain["mu"])
    r = run(built, ["--analysis-type", "ppbayes", "--data-file", base + ".bed", "--pheno", base + ".phen", "--chain-length", str(iters),
                    "--burn-in", "0", "--thin", "1", "--seed", str(int(g.z["seed"])), "--S", ",".join("%.17g" % s for s in g.S[0]),
                    "--mcmc-samples", out, "--sample-precision", "17"] + sargs)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "mean flowgraph duration" in r.stdout
    rows = read_samples(out, g.M, g.N, g.nG)
    assert [d["iteration"] for d in rows] == list(range(iters))

    def close(a, b, what, it, tol=1e-8):
        a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
        err = np.max(np.abs(a - b) / (1.0 + np.abs(b)))
        assert err <= tol, "%s differs at iteration %d: %.3e" % (what, it, err)

    if layout != "dense":
        # the sparse chains of the fixture were recorded after the dense one in ONE process: their marker orders continue the
        # libc rand() stream, a fresh process restarts it.  The scalar draws (mt19937) still match until the orders differ,
        # i.e. for iteration 0's mu only; the full comparison is made for the dense chain.
        close(rows[0]["mu"], chain["mu"][0], "mu", 0)
        return
    full = list(chain["full_iters"])
    for it, d in enumerate(rows):
        close(d["mu"], chain["mu"][it], "mu", it)
        close(d["sigmaE"], chain["sigmaE_post"][it], "sigmaE", it)
        close(d["sigmaG"], chain["sigmaG_post"][it], "sigmaG", it)
        close(d["beta"].sum(), chain["chk_beta"][it], "sum(beta)", it)
        close(d["comp"].sum(), chain["chk_comp"][it], "sum(components)", it)
        close(d["acum"].sum(), chain["chk_acum"][it], "sum(acum)", it)
        close((d["eps"] ** 2).sum(), chain["chk_eps2"][it], "||eps||^2", it)
        if it in full:
            k = full.index(it)
            close(d["beta"], chain["beta_full"][k], "beta", it)
            assert (d["comp"] == chain["comp_full"][k]).all()
            close(d["acum"], chain["acum_full"][k], "acum", it)
            close(d["eps"], chain["eps_full"][k], "eps", it)


@pytest.mark.gpu
def test_async_chain_runs(built, tmp_path):
    g = Golden("missing_small")
    base = write_plink(g, str(tmp_path), groups_file=True)
    out = str(tmp_path / "samples.csv")
    r = run(built, ["--analysis-type", "asyncppbayes", "--analysis-tokens", "20", "--data-file", base + ".bed", "--pheno", base + ".phen",
                    "--group", base + ".group", "--S", ";".join(",".join("%.17g" % s for s in row) for row in g.S), "--chain-length", "20",
                    "--burn-in", "10", "--thin", "5", "--seed", "3", "--mcmc-samples", out, "--no-epsilon"])
    assert r.returncode == 0, r.stderr + r.stdout
    rows = np.loadtxt(out, delimiter=",", skiprows=1, ndmin=2)
    assert rows.shape == (2, 2 + g.M + 1 + g.nG + 2 * g.M) and list(rows[:, 0]) == [10, 15]
    assert np.isfinite(rows).all() and (rows[:, 2 + g.M] > 0).all()


@pytest.mark.gpu
def test_posterior_on_test_gauss_agrees_with_the_reference_chain(built, tmp_path):
    """Tier 3 (BASELINE.json north_star): posterior summaries of a 300-iteration ppbayes chain on test.gauss (burn-in 100) against
    the reference's own chain from the same seed (tests/golden/gauss_long.npz, made by oracle/make_golden.py).  The chains start
    identical; floating-point differences may eventually flip a draw, after which they are two samples of the same posterior,
    so the gate is agreement within Monte Carlo error (a fraction of the posterior sd) -- and it reports how long they stayed
    identical."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "gauss_long.npz"))
    g = Golden("gauss")
    base = write_plink(g, str(tmp_path))
    out = str(tmp_path / "samples.csv")
    iters, burn = int(z["iterations"]), int(z["burn_in"])
    r = run(built, ["--analysis-type", "ppbayes", "--data-file", base + ".bed", "--pheno", base + ".phen", "--chain-length", str(iters),
                    "--burn-in", "0", "--thin", "1", "--seed", str(int(z["seed"])), "--S", ",".join("%.17g" % s for s in g.S[0]),
                    "--mcmc-samples", out, "--sample-precision", "12", "--no-epsilon"])
    assert r.returncode == 0, r.stderr + r.stdout
    M = g.M
    cols = [0, 1, 2 + M, 3 + M]                                  # iteration, mu, sigmaE, sigmaG[1]
    data = np.loadtxt(out, delimiter=",", skiprows=1, usecols=cols, ndmin=2)
    assert data.shape[0] == iters
    mu, sigmaE, sigmaG = data[:, 1], data[:, 2], data[:, 3]
    same = np.abs(sigmaE - z["sigmaE"]) <= 1e-8 * (1 + np.abs(z["sigmaE"]))
    n_same = iters if same.all() else int(np.argmin(same))
    print("chains identical (1e-8) for the first %d of %d iterations" % (n_same, iters))
    assert n_same >= 10
    for name, got, ref in (("sigmaE", sigmaE, z["sigmaE"]), ("sigmaG", sigmaG, z["sigmaG"]), ("mu", mu, z["mu"])):
        gm, rm, sd = got[burn:].mean(), ref[burn:].mean(), max(ref[burn:].std(), 1e-12)
        assert abs(gm - rm) <= 0.5 * sd, "%s: posterior mean %.6g vs reference %.6g (posterior sd %.3g)" % (name, gm, rm, sd)
    # effect sizes: posterior-mean betas of the last 200 iterations
    beta = np.loadtxt(out, delimiter=",", skiprows=1 + burn, usecols=range(2, 2 + M), ndmin=2).mean(axis=0)
    ref = z["beta_mean"]
    assert abs((beta ** 2).sum() - (ref ** 2).sum()) <= 0.5 * max((ref ** 2).sum(), 1e-12) + 1e-6
    top = np.argsort(-np.abs(ref))[:20]
    assert np.corrcoef(beta[top], ref[top])[0, 1] > 0.5 or n_same == iters


@pytest.mark.parametrize("sparse", [None, "ragged", "eigen"])
@pytest.mark.parametrize("compress", [False, True])
def test_preprocess_writes_the_references_files(built, tmp_path, sparse, compress):
    """`brr_gpu --preprocess [--compress] [--sparse-data ...]` (host work only, no GPU): the files equal what the REFERENCE's own markers
    and index writer produce (oracle/_ref: Marker::write / compress, src/preprocessgraph.cpp:92-125) -- byte for byte for the index-list
    layouts, to 1e-13 for the dense columns (the column mean is summed in another order)."""
    import ref_lib as R
    if not R.ref_available():
        pytest.skip("oracle/_ref is not built")
    g = Golden("missing_small")
    base = write_plink(g, str(tmp_path))
    layout, ext = {None: (0, ".ppbed"), "eigen": (2, ".eigen.sparsebed"), "ragged": (1, ".ragged.sparsebed")}[sparse]
    r = run(built, ["--preprocess", "--data-file", base + ".bed"] + (["--compress"] if compress else []) + (["--sparse-data", sparse] if sparse else []))
    assert r.returncode == 0, r.stderr
    ref = str(tmp_path / "ref")
    R.write_pp_files(g.bed, g.N, g.M, layout, int(compress), ref + ext, ref + ext + "index")
    own_idx = np.fromfile(base + ext + "index", dtype=np.uint64).reshape(g.M, 3); ref_idx = np.fromfile(ref + ext + "index", dtype=np.uint64).reshape(g.M, 3)
    assert np.array_equal(own_idx[:, 2], ref_idx[:, 2])                 # original sizes
    if sparse is not None:
        assert np.array_equal(own_idx, ref_idx)
        assert open(base + ext, "rb").read() == open(ref + ext, "rb").read()
    elif not compress:
        assert np.array_equal(own_idx, ref_idx)
        assert np.max(np.abs(np.fromfile(base + ext) - np.fromfile(ref + ext))) <= 1e-13
    else:                                                               # deflated doubles that differ in the last bit: compare after inflation
        import zlib
        blob = open(base + ext, "rb").read(); rblob = open(ref + ext, "rb").read()
        for j in (0, g.M // 2, g.M - 1):
            a = np.frombuffer(zlib.decompress(blob[int(own_idx[j, 0]):int(own_idx[j, 0] + own_idx[j, 1])]), dtype=np.float64)
            b = np.frombuffer(zlib.decompress(rblob[int(ref_idx[j, 0]):int(ref_idx[j, 0] + ref_idx[j, 1])]), dtype=np.float64)
            assert a.size == g.N and np.max(np.abs(a - b)) <= 1e-13


@pytest.mark.gpu
def test_iteration_and_column_logs(built, tmp_path):
    """--iterLog / --colLog (src/logwriter.cpp, src/BayesRBase.cpp:261-273): the iteration log has the reference's header and one line per
    iteration whose sigmaG / sigmaE / mu are the chain's; the column log has one line per sweep."""
    g = Golden("uk10k")
    base = write_plink(g, str(tmp_path))
    out, ilog, clog = str(tmp_path / "s.csv"), str(tmp_path / "iter.csv"), str(tmp_path / "col.csv")
    r = run(built, ["--analysis-type", "ppbayes", "--data-file", base + ".bed", "--pheno", base + ".phen", "--chain-length", "3", "--burn-in", "0",
                    "--thin", "1", "--seed", "3", "--S", ",".join("%.17g" % x for x in g.S[0]), "--mcmc-samples", out, "--sample-precision", "17",
                    "--iterLog", ilog, "--colLog", clog, "--out", "run1", "--no-epsilon"])
    assert r.returncode == 0, r.stderr + r.stdout
    lines = open(ilog).read().strip().splitlines()
    assert lines[0] == "iter,m_0,sigmaG,sigmaE,mu,mu_upd_microS,sigmaG_upd_microS,sigmaE_upd_microS,flowgraph_milliS,iter_d_milliS"
    log = np.loadtxt(ilog, delimiter=",", skiprows=1, ndmin=2)
    assert log.shape == (3, 10) and list(log[:, 0]) == [0, 1, 2]
    rows = np.loadtxt(out, delimiter=",", skiprows=1, ndmin=2)
    assert np.allclose(log[:, 4], rows[:, 1], rtol=1e-5)               # mu (the log is written with the stream's default precision)
    assert np.allclose(log[:, 3], rows[:, 2 + g.M], rtol=1e-5)         # sigmaE
    assert np.allclose(log[:, 2], rows[:, 3 + g.M], rtol=1e-5)         # sigmaG
    col = np.loadtxt(clog, delimiter=",", skiprows=1, ndmin=2)
    assert col.shape == (3, 4) and (col[:, 1] == g.M).all() and (col[:, 3] > 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("sparse", [None, "ragged"])
def test_marker_cache_run_from_the_references_preprocessed_files(built, tmp_path, sparse):
    """`--marker-cache --compress`: the chain fed from files the REFERENCE's markers wrote (no .bed on disk) equals the .bed-fed chain."""
    import ref_lib as R
    if not R.ref_available():
        pytest.skip("oracle/_ref is not built")
    g = Golden("missing_small")
    base = write_plink(g, str(tmp_path))
    layout, ext = (0, ".ppbed") if sparse is None else (1, ".ragged.sparsebed")
    R.write_pp_files(g.bed, g.N, g.M, layout, 1, base + ext, base + ext + "index")
    S = ";".join(",".join("%.9g" % x for x in row) for row in g.S)
    common = ["--data-file", base + ".bed", "--pheno", base + ".phen", "--analysis-type", "ppbayes", "--chain-length", "4", "--burn-in", "0",
              "--thin", "1", "--seed", "7", "--S", S, "--sample-precision", "17"] + (["--sparse-data", sparse] if sparse else [])
    if g.groups is not None:
        write_plink(g, str(tmp_path), groups_file=True)
        common += ["--group", base + ".group"]
    r1 = run(built, common + ["--mcmc-samples", base + "_bed.csv"])
    assert r1.returncode == 0, r1.stderr
    os.rename(base + ".bed", base + ".bed.away")
    r2 = run(built, common + ["--marker-cache", "--compress", "--mcmc-samples", base + "_pp.csv"])
    assert r2.returncode == 0, r2.stderr
    a = np.loadtxt(base + "_bed.csv", delimiter=",", skiprows=1); b = np.loadtxt(base + "_pp.csv", delimiter=",", skiprows=1)
    assert a.shape == b.shape and np.max(np.abs(a - b) / (1 + np.abs(a))) <= 1e-8


@pytest.mark.gpu
def test_fixed_effects_chain_from_seed_reproduces_the_reference(built, tmp_path):
    """--fixed_effects / --fixedEffectNumber (src/BayesRBase.cpp:214-235): the reference's own runGibbs with two covariates (oracle/_ref,
    run here) against the host's chain from the same seed: mu, beta, sigmaE, sigmaG, gamma per iteration, the residual at the end."""
    import ref_lib as R
    if not R.ref_available():
        pytest.skip("oracle/_ref is not built")
    g = Golden("missing_small")
    rng = np.random.default_rng(12)
    X = np.c_[rng.normal(size=g.N), (rng.random(g.N) < 0.4).astype(np.float64)].astype(np.float32)
    y = (np.asarray(g.y, dtype=np.float64) + 0.8 * X[:, 0] - 0.5 * X[:, 1]).astype(np.float32)
    iters, seed = 5, 11
    arr, _ = R.run_chain(g.bed, g.N, g.M, y, g.S, seed=seed, iterations=iters, layout=0, groups=g.groups, X=X)
    g.y = y
    base = write_plink(g, str(tmp_path), groups_file=g.groups is not None)
    with open(base + ".cov", "w") as f:
        for i in range(g.N):
            f.write(",".join("%.9g" % v for v in X[i]) + "\n")
    out = str(tmp_path / "samples.csv")
    S = ";".join(",".join("%.17g" % x for x in row) for row in g.S)
    args = ["--analysis-type", "ppbayes", "--data-file", base + ".bed", "--pheno", base + ".phen", "--chain-length", str(iters), "--burn-in", "0",
            "--thin", "1", "--seed", str(seed), "--S", S, "--mcmc-samples", out, "--sample-precision", "17",
            "--fixed_effects", base + ".cov", "--fixedEffectNumber", "2"] + (["--group", base + ".group"] if g.groups is not None else [])
    r = run(built, args)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "gamma[2]" in open(out).readline()
    rows = read_samples(out, g.M, g.N, g.nG, nF=2)

    def close(a, b, what, it, tol=1e-8):
        err = np.max(np.abs(np.asarray(a, float) - np.asarray(b, float)) / (1.0 + np.abs(np.asarray(b, float))))
        assert err <= tol, "%s differs at iteration %d: %.3e" % (what, it, err)

    for it, d in enumerate(rows):
        close(d["mu"], arr["mu"][it], "mu", it)
        close(d["beta"], arr["beta_out"][it], "beta", it)
        close(d["sigmaE"], arr["sigmaE_post"][it], "sigmaE", it)
        close(d["sigmaG"], arr["sigmaG_post"][it], "sigmaG", it)
        close(d["gamma"], arr["gamma_post"][it], "gamma", it)
    close(rows[-1]["eps"], arr["eps_post"][-1], "eps", iters - 1)
    assert np.abs(arr["gamma_post"][-1]).max() > 0.1            # the covariates carry signal: the effects are not trivially zero

#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY: generate tests/golden/*.npz from the UNMODIFIED reference.

Runs here (needs /root/reference and oracle/_ref/libbrr_ref.so, built by `make -C oracle ref`);
the outputs are committed so that the tests, smoke() and bench.py never read /root/reference.

    python oracle/make_golden.py

Every chain is the reference's own BayesRBase::runGibbs (src/BayesRBase.cpp:130-285) with the
std-backed RNG shim (oracle/shim/boost), driven through the AnalysisGraph interface with a tap
(oracle/ref_harness.cpp).  Per iteration we record what AnalysisGraph::exec consumes (marker order,
the (u,z) table, pi, sigmaG, sigmaE, mu) and what it must leave behind (beta, components, acum,
epsilon, v, betasqnG, epsilonSum).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_lib as R  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
FULL_AT = (0, 9)      # iterations whose full beta/comp/acum vectors are stored


def chain_record(bed, N, M, y, S, seed, iters, layout, groups=None, full_at=FULL_AT, all_full=False):
    arr, _ = R.run_chain(bed, N, M, y, S, seed=seed, iterations=iters, layout=layout, groups=groups)
    rec = dict(
        mu=arr["mu"], sigmaE_in=arr["sigmaE_in"], sigmaG_in=arr["sigmaG_in"], pi_in=arr["pi_in"],
        sigmaE_post=arr["sigmaE_post"], sigmaG_post=arr["sigmaG_post"], pi_post=arr["pi_post"],
        order=arr["order"].astype(np.uint16 if M < 65536 else np.uint32), rand_uz=arr["rand_uz"],
        eps0=arr["eps_in"][0], epssum_in=arr["epssum_in"], epssum_out=arr["epssum_out"],
        v_out=arr["v_out"], betasq_out=arr["betasq_out"],
        chk_beta=arr["beta_out"].sum(1), chk_beta2=(arr["beta_out"] ** 2).sum(1),
        chk_comp=arr["comp_out"].sum(1), chk_acum=arr["acum_out"].sum(1),
        chk_eps2=(arr["eps_out"] ** 2).sum(1), chk_eps=arr["eps_out"].sum(1),
        nonskipped=((arr["beta_in"] != 0) | (arr["beta_out"] != 0)).sum(1).astype(np.int64),
    )
    its = range(iters) if all_full else [i for i in full_at if i < iters]
    rec["full_iters"] = np.asarray(list(its), dtype=np.int64)
    rec["beta_full"] = arr["beta_out"][list(its)]
    rec["comp_full"] = arr["comp_out"][list(its)].astype(np.uint8)
    rec["acum_full"] = arr["acum_out"][list(its)]
    rec["eps_full"] = arr["eps_out"][list(its)]
    return rec


def tier1_record(bed, N, cols, eps, rng):
    """computeNum / epsilon change of the reference's three kernels on chosen columns."""
    out = {k: [] for k in ("num_dense", "num_ragged", "num_eigen", "beta_old", "beta_new", "epssum_delta_ragged",
                           "epssum_delta_eigen", "lut", "ragged_stats", "eigen_stats")}
    deltas = {"dense": [], "ragged": [], "eigen": []}
    eps_sum = float(eps.sum())
    for n, j in enumerate(cols):
        b0 = 0.0 if n % 3 == 0 else float(rng.normal() * 0.01)
        b1 = float(rng.normal() * 0.01)
        nd, dd, _ = R.marker_ops(0, bed[j], N, eps, b0, b1)
        nr, dr, er = R.marker_ops(1, bed[j], N, eps, b0, b1, eps_sum)
        ne, de, ee = R.marker_ops(2, bed[j], N, eps, b0, b1, eps_sum)
        out["num_dense"].append(nd); out["num_ragged"].append(nr); out["num_eigen"].append(ne)
        out["beta_old"].append(b0); out["beta_new"].append(b1)
        out["epssum_delta_ragged"].append(er); out["epssum_delta_eigen"].append(ee)
        cx = R.dense_column(bed[j], N)
        cls = decode(bed[j], N)
        lut = np.full(4, np.nan)
        for c in range(4):
            if (cls == c).any():
                lut[c] = cx[cls == c][0]
                assert (cx[cls == c] == lut[c]).all()      # <=4 distinct values per dense column
        out["lut"].append(lut)
        out["ragged_stats"].append(R.ragged_marker(bed[j], N)[0])
        out["eigen_stats"].append(R.eigen_stats(bed[j], N))
        if n < 8:
            deltas["dense"].append(dd); deltas["ragged"].append(dr); deltas["eigen"].append(de)
    rec = {k: np.asarray(v) for k, v in out.items()}
    rec["cols"] = np.asarray(cols, dtype=np.int64)
    rec["eps"] = eps
    for k, v in deltas.items():
        rec["delta_" + k] = np.asarray(v)
    return rec


def decode(bedcol, N):
    i = np.arange(N)
    two = (bedcol[i >> 2] >> ((i & 3) * 2)) & 3
    a1 = 1 - (two & 1); a2 = 1 - ((two >> 1) & 1)
    return np.where((a1 == 0) & (a2 == 1), 3, a1 + a2).astype(np.uint8)


def standardise(y):
    yd = y.astype(np.float64)
    yd = yd - yd.mean()
    return yd / np.sqrt((yd ** 2).sum() / (len(yd) - 1))


def share_inputs(base, other):
    """Drop the marker order / (u,z) table of `other` when identical to `base` (same RNG stream)."""
    if (other["order"] == base["order"]).all() and (other["rand_uz"] == base["rand_uz"]).all():
        other = dict(other); del other["order"], other["rand_uz"]
        other["inputs_shared_with_dense"] = np.int64(1)
    return other


def save(name, **kw):
    if "dense" in kw:
        for k in ("ragged", "eigen"):
            if k in kw:
                kw[k] = share_inputs(kw["dense"], kw[k])
    flat = {}
    for k, v in kw.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                flat[k + "__" + kk] = vv
        else:
            flat[k] = v
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **flat)
    print(name, "%.2f MB" % (os.path.getsize(path) / 1e6))


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20260101)

    # S1 / config 0: test/data uk10k, --S 0.01,0.001,0.0001, 10 iterations, seed 1
    d = REF + "/test/data/"
    ids = R.read_fam(d + "uk10k_chr1_1mb.fam"); N = len(ids); M = R.read_bim_count(d + "uk10k_chr1_1mb.bim")
    bed = R.read_bed(d + "uk10k_chr1_1mb.bed", N, M)
    y = R.read_phen(d + "test.phen", ids)
    S = np.array([[0.01, 0.001, 0.0001]])
    cols = np.unique(np.concatenate([np.arange(0, M, 13), [M - 1]]))
    save("uk10k", bed=bed, y=y, S=S, N=N, M=M, seed=1,
         dense=chain_record(bed, N, M, y, S, 1, 10, 0),
         ragged=chain_record(bed, N, M, y, S, 1, 10, 1, full_at=(9,)),
         eigen=chain_record(bed, N, M, y, S, 1, 10, 2, full_at=(9,)),
         tier1=tier1_record(bed, N, cols, standardise(y), rng))

    # S5: test.groups (same genotypes, 2 groups, S from test.cva)
    g = REF + "/test.groups/"
    yg = R.read_phen(g + "test.phen", R.read_fam(g + "uk10k_chr1_1mb.fam"))
    groups = R.read_groups(g + "test.group")
    Sg = np.array([[0.001, 0.01], [0.001, 0.01]])
    assert (R.read_bed(g + "uk10k_chr1_1mb.bed", N, M) == bed).all()
    save("groups", y=yg, S=Sg, groups=groups, N=N, M=M, seed=1, bed_from="uk10k",
         dense=chain_record(bed, N, M, yg, Sg, 1, 10, 0, groups=groups),
         ragged=chain_record(bed, N, M, yg, Sg, 1, 10, 1, groups=groups, full_at=(9,)))

    # S2: test.gauss
    d = REF + "/test.gauss/"
    ids = R.read_fam(d + "data.fam"); N2 = len(ids); M2 = R.read_bim_count(d + "data.bim")
    bed2 = R.read_bed(d + "data.bed", N2, M2)
    y2 = R.read_phen(d + "data.phen", ids)
    cols2 = np.arange(0, M2, 37)
    # S2 long: 300 iterations of the reference chain on test.gauss, posterior summaries only (tier 3: agreement within Monte Carlo error)
    arr, _ = R.run_chain(bed2, N2, M2, y2, S, seed=11, iterations=300, layout=0)
    np.savez_compressed(os.path.join(OUT, "gauss_long.npz"), N=N2, M=M2, S=S, seed=11, iterations=300, burn_in=100, bed_from="gauss", y=y2,
                        mu=arr["mu"], sigmaE=arr["sigmaE_post"], sigmaG=arr["sigmaG_post"][:, 0], m0=arr["v_out"][:, 0, 1:].sum(axis=1),
                        beta2=(arr["beta_out"] ** 2).sum(axis=1), beta_mean=arr["beta_out"][100:].mean(axis=0),
                        pip=(arr["comp_out"][100:] != 0).mean(axis=0).astype(np.float32))
    del arr
    save("gauss", bed=bed2, y=y2, S=S, N=N2, M=M2, seed=1,
         dense=chain_record(bed2, N2, M2, y2, S, 1, 10, 0),
         tier1=tier1_record(bed2, N2, cols2, standardise(y2), rng))

    # Edge cases the fixtures do not hold (zero missing genotypes upstream): ragged N
    # (not a multiple of 4 / 32), missing genotypes, a monomorphic-but-one column, 2 groups, K=3.
    N3, M3 = 1003, 257
    maf = rng.uniform(0.02, 0.5, M3)
    geno = rng.binomial(2, maf[:, None], size=(M3, N3)).astype(np.uint8)
    miss = rng.random((M3, N3)) < 0.03
    geno[5] = 0; geno[5, 17] = 1                       # nearly monomorphic
    miss[6] = False                                    # a column without missing
    code = np.where(miss, 1, np.array([3, 2, 0], dtype=np.uint8)[geno])   # PLINK: 00->2 10->1 11->0 01->missing
    bed3 = np.zeros((M3, (N3 + 3) // 4), dtype=np.uint8)
    for k in range(4):
        sub = code[:, k::4]
        bed3[:, :sub.shape[1]] |= (sub << (2 * k)).astype(np.uint8)
    beta_true = np.zeros(M3); beta_true[rng.choice(M3, 20, replace=False)] = rng.normal(size=20)
    y3 = ((geno * ~miss).T @ beta_true + rng.normal(size=N3) * 2).astype(np.float32)
    groups3 = (np.arange(M3) % 2).astype(np.int32)
    S3 = np.array([[0.001, 0.01], [0.0005, 0.02]])
    save("missing_small", bed=bed3, y=y3, S=S3, groups=groups3, N=N3, M=M3, seed=7,
         dense=chain_record(bed3, N3, M3, y3, S3, 7, 6, 0, groups=groups3, all_full=True),
         ragged=chain_record(bed3, N3, M3, y3, S3, 7, 6, 1, groups=groups3, all_full=True),
         eigen=chain_record(bed3, N3, M3, y3, S3, 7, 6, 2, groups=groups3, all_full=True),
         tier1=tier1_record(bed3, N3, np.arange(M3), standardise(y3), rng))


if __name__ == "__main__":
    main()

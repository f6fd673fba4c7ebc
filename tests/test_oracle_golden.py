"""CPU: the plain-C restatement (oracle/bayesr_oracle.c) against the golden vectors the UNMODIFIED
reference produced (oracle/make_golden.py).  This pins the oracle the GPU tests are checked against."""
import numpy as np
import pytest

import oracle_lib as O
from helpers import Golden, check_iteration, num_tol, replay


@pytest.fixture(scope="module", autouse=True)
def _build():
    O.build()


def _replay_oracle(g, layout, iters, tokens=0):
    chain = g.chain(layout)
    st = dict(eps=None, beta=np.zeros(g.M), comp=np.zeros(g.M), acum=np.zeros(g.M))

    def sweep(it, order, rand_uz, pi, sigmaG, sigmaE, eps_sum):
        mdl = O.ModelArrays(g.N, g.S, pi, sigmaG, sigmaE)
        return O.sweep(layout, g.bed, g.M, order, rand_uz, mdl, g.groups,
                       st["eps"], st["beta"], st["comp"], st["acum"], tokens=tokens, eps_sum=eps_sum)

    def get_state():
        return st["eps"].copy(), st["beta"].copy(), st["comp"].copy(), st["acum"].copy()

    def set_eps(e):
        st["eps"] = np.ascontiguousarray(e, dtype=np.float64).copy()

    return chain, replay(chain, iters, sweep, get_state, set_eps)


@pytest.mark.parametrize("name,layout,iters", [("uk10k", "dense", 3), ("uk10k", "ragged", 3), ("groups", "dense", 2),
                                               ("gauss", "dense", 2), ("missing_small", "dense", 6),
                                               ("missing_small", "ragged", 6), ("missing_small", "eigen", 6), ("uk10k", "eigen", 2)])
def test_oracle_replays_reference_chain(name, layout, iters):
    g = Golden(name)
    chain, it_res = _replay_oracle(g, layout, iters)
    for it, res in it_res:
        check_iteration(chain, it, res, tol=1e-9)
        if layout != "dense":
            assert abs(res["eps_sum"] - chain["epssum_out"][it]) <= 1e-9 * (1 + abs(chain["epssum_out"][it]))


@pytest.mark.parametrize("name", ["uk10k", "gauss", "missing_small"])
def test_oracle_tier1_matches_reference_kernels(name):
    g = Golden(name)
    t = g.tier1()
    eps = t["eps"]; en = np.linalg.norm(eps); es = float(eps.sum())
    for n, j in enumerate(t["cols"]):
        col = g.bed[j]
        cx = O.dense_column(col, g.N)
        cls = O.decode(col, g.N)
        lut = O.dense_lut(col, g.N)
        for c in range(4):
            if not np.isnan(t["lut"][n][c]):
                assert abs(lut[c] - t["lut"][n][c]) <= 1e-12 * (1 + abs(lut[c]))
        b0, b1 = float(t["beta_old"][n]), float(t["beta_new"][n])
        tol = num_tol(t["num_dense"][n], np.linalg.norm(cx), en)
        assert abs(O.dense_compute_num(cx, eps, b0) - t["num_dense"][n]) <= tol
        stats, ones, twos, miss = O.ragged_marker(col, g.N)
        assert np.allclose(stats, t["ragged_stats"][n], rtol=1e-14, atol=0)
        assert abs(O.ragged_compute_num(stats, ones, twos, eps, b0, es) - t["num_ragged"][n]) <= tol
        est = O.eigen_stats(col, g.N)
        assert np.allclose(est, t["eigen_stats"][n], rtol=1e-14, atol=0)
        assert abs(O.eigen_compute_num(est, cls, eps, b0, es) - t["num_eigen"][n]) <= tol
        if n < len(t["delta_dense"]):
            e = eps.copy(); O.dense_eps_update(cx, e, b0, b1)
            assert np.max(np.abs((e - eps) - t["delta_dense"][n])) <= 1e-12
            e = eps.copy(); d = O.ragged_eps_update(stats, ones, twos, miss, e, b0, b1, 0.0)
            assert np.max(np.abs((e - eps) - t["delta_ragged"][n])) <= 1e-12
            assert abs(d - t["epssum_delta_ragged"][n]) <= 1e-12 * (1 + abs(d))
            e = eps.copy(); d = O.eigen_eps_update(est, cls, e, b0, b1, 0.0)
            assert np.max(np.abs((e - eps) - t["delta_eigen"][n])) <= 1e-12
            assert abs(d - t["epssum_delta_eigen"][n]) <= 1e-12 * (1 + abs(d))


def test_no_missing_layouts_agree():
    """without missing genotypes dense and ragged give the same chain."""
    g = Golden("uk10k")
    d, r = g.chain("dense"), g.chain("ragged")
    assert np.allclose(d["chk_beta2"], r["chk_beta2"], rtol=1e-9)
    assert (d["chk_comp"] == r["chk_comp"]).all()


def test_async_oracle_with_one_token_is_the_sync_chain():
    g = Golden("missing_small")
    chain, it_res = _replay_oracle(g, "dense", 2, tokens=1)
    for it, res in it_res:
        check_iteration(chain, it, res, tol=1e-9, check_acum=False)

"""GPU parity tests, ragged and eigen-sparse layouts (index-list kernels), through the C ABI."""
import numpy as np
import pytest

import oracle_lib as O
from bayesrrcmd_b200 import Engine
from helpers import Golden, check_iteration, num_tol, replay
from test_gpu_dense import _random_problem

pytestmark = pytest.mark.gpu


def make_engine(g, layout, **kw):
    e = Engine(g.N, g.M, g.K, g.nG, layout=layout, **kw)
    e.upload_bed(g.bed)
    e.finalize()
    e.set_model(g.S, g.groups)
    return e


@pytest.mark.parametrize("name,layout", [("uk10k", "ragged"), ("missing_small", "ragged"), ("missing_small", "eigen"),
                                         ("gauss", "eigen")])
def test_tier1_sparse_kernels(name, layout):
    g = Golden(name)
    t = g.tier1()
    e = make_engine(g, layout)
    stats = e.ragged_stats()
    eps = t["eps"]; en = np.linalg.norm(eps); es = float(eps.sum())
    e.set_state(eps=eps)
    key = "num_" + layout
    for n, j in enumerate(t["cols"]):
        ref_stats = t[layout + "_stats"][n]
        assert np.allclose(stats[j], ref_stats, rtol=1e-13, atol=0), (j, stats[j], ref_stats)
        num = e.compute_num(int(j), float(t["beta_old"][n]), es)
        assert abs(num - t[key][n]) <= num_tol(t[key][n], np.sqrt(g.N - 1.0), en), (j, num, t[key][n])
    for n in range(len(t["delta_" + layout])):
        j = int(t["cols"][n])
        e.set_state(eps=eps)
        d = e.apply_delta(j, float(t["beta_old"][n]), float(t["beta_new"][n]))
        got = e.get_state(beta=False, comp=False, acum=False)[0] - eps
        assert np.max(np.abs(got - t["delta_" + layout][n])) <= 1e-12
        assert abs(d - t["epssum_delta_" + layout][n]) <= 1e-11 * (1 + abs(d))


@pytest.mark.parametrize("name,layout,window", [("uk10k", "ragged", 16), ("uk10k", "eigen", 8), ("groups", "ragged", 32),
                                                ("missing_small", "ragged", 16), ("missing_small", "eigen", 4),
                                                ("missing_small", "ragged", 1)])
def test_tier2_replay_reference_chain(name, layout, window):
    g = Golden(name)
    e = make_engine(g, layout, block_markers=window)
    chain = g.chain(layout)
    e.set_state(eps=np.zeros(g.N), beta=np.zeros(g.M), comp=np.zeros(g.M), acum=np.zeros(g.M))
    outs = []

    def sweep(it, order, rand_uz, pi, sigmaG, sigmaE, eps_sum):
        v, bs, out = e.sweep(order, rand_uz, pi, sigmaG, sigmaE, eps_sum, 0)
        outs.append(out)
        return v, bs, out.eps_sum

    for it, res in replay(chain, len(chain["mu"]), sweep, e.get_state, lambda x: e.set_state(eps=x)):
        check_iteration(chain, it, res, tol=1e-8)
        assert outs[-1].n_nonskipped == chain["nonskipped"][it]
        assert abs(res["eps_sum"] - chain["epssum_out"][it]) <= 1e-8 * (1 + abs(chain["epssum_out"][it]))


@pytest.mark.parametrize("layout,N,M,p_miss,nG,K,window,tokens", [
    ("ragged", 1000, 300, 0.01, 2, 3, 16, 0),
    ("eigen", 5003, 200, 0.02, 1, 5, 8, 0),
    ("ragged", 70001, 64, 0.001, 1, 4, 16, 0),
    ("ragged", 3000, 257, 0.01, 2, 4, 16, 20),
    ("eigen", 3000, 130, 0.05, 1, 4, 16, 7),
])
def test_sweep_matches_oracle_on_random_inputs(layout, N, M, p_miss, nG, K, window, tokens):
    O.build()
    pr = _random_problem(N * 3 + M, N, M, p_miss, nG, K)
    rng = pr["rng"]
    e = Engine(N, M, K, nG, layout=layout, block_markers=window)
    e.upload_bed(pr["bed"]); e.finalize(); e.set_model(pr["S"], pr["groups"])
    beta = np.where(rng.random(M) < 0.3, rng.normal(size=M) * 0.05, 0.0)
    comp = (beta != 0).astype(np.float64); acum = rng.random(M)
    eps_o = pr["eps"].copy(); beta_o = beta.copy(); comp_o = comp.copy(); acum_o = acum.copy()
    e.set_state(eps=pr["eps"], beta=beta, comp=comp, acum=acum)
    es = float(pr["eps"].sum())
    for sweep_i in range(2):
        order = rng.permutation(M).astype(np.uint32)
        rand_uz = np.c_[rng.random(M), rng.normal(size=M)]
        mdl = O.ModelArrays(N, pr["S"], pr["pi"], pr["sigmaG"], pr["sigmaE"])
        v_o, bs_o, es_o = O.sweep(layout, pr["bed"], M, order, rand_uz, mdl, pr["groups"], eps_o, beta_o, comp_o, acum_o,
                                  tokens=tokens, eps_sum=es)
        v, bs, out = e.sweep(order, rand_uz, pr["pi"], pr["sigmaG"], pr["sigmaE"], es, tokens)
        eps, b, c, a = e.get_state()
        assert (c == comp_o).all() and (v == v_o).all()
        assert np.max(np.abs(b - beta_o)) <= 1e-9
        assert np.max(np.abs(a - acum_o)) <= 1e-9
        assert np.max(np.abs(eps - eps_o)) <= 1e-9
        assert np.max(np.abs(bs - bs_o)) <= 1e-9 * (1 + np.abs(bs_o).max())
        assert abs(out.eps_sum - es_o) <= 1e-9 * (1 + abs(es_o))
        es = es_o


def test_upload_lists_directly_and_reject_bad_lists():
    from bayesrrcmd_b200 import BrrGpuError
    N, M = 900, 3
    ones = [np.array([1, 5, 700], np.int32), np.array([], np.int32), np.array([0, 899], np.int32)]
    twos = [np.array([2], np.int32), np.array([10, 11], np.int32), np.array([], np.int32)]
    miss = [np.array([], np.int32), np.array([500], np.int32), np.array([], np.int32)]
    idx = np.concatenate([x for j in range(M) for x in (ones[j], twos[j], miss[j])])
    off = np.cumsum([0] + [len(x) for j in range(M) for x in (ones[j], twos[j], miss[j])]).astype(np.uint64)
    e = Engine(N, M, 4, 1, layout="ragged")
    e.upload_ragged(off, idx); e.finalize()
    assert (e.counts()[:, 1:] == np.array([[3, 1, 0], [0, 2, 1], [2, 0, 0]])).all()
    eps = np.random.default_rng(1).normal(size=N)
    e.set_state(eps=eps)
    st = e.ragged_stats()
    num = e.compute_num(0, 0.0, float(eps.sum()))
    ref = -st[0, 0] * eps.sum() / st[0, 1] + (eps[[1, 5, 700]].sum() + 2 * eps[2]) / st[0, 1]
    assert abs(num - ref) <= 1e-12 * (1 + abs(ref))
    bad = idx.copy(); bad[0], bad[1] = bad[1], bad[0]
    with pytest.raises(BrrGpuError, match="ascending"):
        e.upload_ragged(off, bad)


def test_sparse_sweep_after_dense_contexts_were_freed():
    """The error word of a sweep must be cleared by the launcher, not inherited from recycled device memory (a sparse sweep
    right after dense contexts were destroyed used to report stale error flags)."""
    import gc
    g = Golden("missing_small")
    chain = g.chain("dense")
    for _ in range(3):
        e = Engine(g.N, g.M, g.K, g.nG, layout="dense")
        e.upload_bed(g.bed); e.finalize(); e.set_model(g.S, g.groups)
        e.set_state(eps=chain["eps0"], beta=np.zeros(g.M), comp=np.zeros(g.M), acum=np.zeros(g.M))
        e.sweep(chain["order"][0].astype(np.uint32), chain["rand_uz"][0], chain["pi_in"][0], chain["sigmaG_in"][0], float(chain["sigmaE_in"][0]))
        del e
        gc.collect()
    for layout in ("ragged", "eigen"):
        e = Engine(g.N, g.M, g.K, g.nG, layout=layout)
        e.upload_bed(g.bed); e.finalize(); e.set_model(g.S, g.groups)
        e.set_state(eps=chain["eps0"], beta=np.zeros(g.M), comp=np.zeros(g.M), acum=np.zeros(g.M))
        v, bs, out = e.sweep(chain["order"][0].astype(np.uint32), chain["rand_uz"][0], chain["pi_in"][0], chain["sigmaG_in"][0],
                             float(chain["sigmaE_in"][0]), float(chain["eps0"].sum()))
        assert v.sum() == g.M


@pytest.mark.parametrize("layout", ["ragged", "eigen"])
def test_device_synthetic_sparse_store_matches_its_host_twin(layout):
    """brr_gpu_synth_ragged (index lists built on the device) must give the very store that the host twin's PLINK payload gives
    through brr_gpu_upload_ragged_from_bed: same counts, same statistics, bit-identical sweeps."""
    from bayesrrcmd_b200.capi import synth_bed_host
    N, M, K = 5000, 96, 4
    S = np.array([[0.01, 0.001, 0.0001]])
    a = Engine(N, M, K, 1, layout=layout); a.synth_ragged(42, 0.001, 0.5, True, 0.01); a.finalize(); a.set_model(S)
    bed = synth_bed_host(42, N, 0, M, maf_lo=0.001, maf_hi=0.5, log_uniform=True, p_missing=0.01)
    b = Engine(N, M, K, 1, layout=layout); b.upload_bed(bed); b.finalize(); b.set_model(S)
    assert (a.counts() == b.counts()).all() and a.counts()[:, 3].sum() > 0
    assert np.array_equal(a.ragged_stats(), b.ragged_stats())
    rng = np.random.default_rng(1)
    eps = rng.normal(size=N)
    pi = np.array([[0.9, 0.05, 0.03, 0.02]]); sigmaG = np.array([0.3])
    for e in (a, b):
        e.set_state(eps=eps, beta=np.zeros(M), comp=np.zeros(M), acum=np.zeros(M))
    es_a = es_b = float(eps.sum())
    for s in range(2):
        order = rng.permutation(M).astype(np.uint32); uz = np.c_[rng.random(M), rng.normal(size=M)]
        va, bsa, oa = a.sweep(order, uz, pi, sigmaG, 0.7, es_a)
        vb, bsb, ob = b.sweep(order, uz, pi, sigmaG, 0.7, es_b)
        es_a, es_b = oa.eps_sum, ob.eps_sum
        for x, y in zip(a.get_state(), b.get_state()):
            assert np.array_equal(x, y)
        assert np.array_equal(va, vb) and oa.n_nonskipped == ob.n_nonskipped

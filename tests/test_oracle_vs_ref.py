"""CPU, only where oracle/_ref was built (this container): the C restatement against the live,
unmodified reference object code on fresh random inputs."""
import numpy as np
import pytest

import oracle_lib as O
import ref_lib as R

pytestmark = pytest.mark.skipif(not R.ref_available(), reason="oracle/_ref not built (needs /root/reference)")


def _random_bed(rng, N, M, p_miss):
    maf = rng.uniform(0.02, 0.5, M)
    geno = rng.binomial(2, maf[:, None], size=(M, N))
    code = np.where(rng.random((M, N)) < p_miss, 1, np.array([3, 2, 0])[geno]).astype(np.uint8)
    bed = np.zeros((M, (N + 3) // 4), dtype=np.uint8)
    for k in range(4):
        sub = code[:, k::4]
        bed[:, :sub.shape[1]] |= (sub << (2 * k)).astype(np.uint8)
    return bed


@pytest.mark.parametrize("N,p_miss", [(257, 0.0), (1001, 0.05)])
def test_marker_ops_match_reference(N, p_miss):
    O.build()
    rng = np.random.default_rng(N)
    bed = _random_bed(rng, N, 12, p_miss)
    eps = rng.normal(size=N)
    for j in range(12):
        b0, b1 = (0.0 if j % 2 else rng.normal() * 0.1), rng.normal() * 0.1
        cx = O.dense_column(bed[j], N)
        assert np.max(np.abs(cx - R.dense_column(bed[j], N))) <= 1e-13
        num, delta, _ = R.marker_ops(0, bed[j], N, eps, b0, b1)
        assert abs(O.dense_compute_num(cx, eps, b0) - num) <= 1e-10 * abs(num) + 1e-12
        stats, ones, twos, miss = O.ragged_marker(bed[j], N)
        rs, ro, rt, rm = R.ragged_marker(bed[j], N)
        assert np.allclose(stats, rs, rtol=1e-15, atol=0) and (ones == ro).all() and (twos == rt).all() and (miss == rm).all()
        num, delta, es = R.marker_ops(1, bed[j], N, eps, b0, b1, float(eps.sum()))
        assert abs(O.ragged_compute_num(stats, ones, twos, eps, b0, float(eps.sum())) - num) <= 1e-10 * abs(num) + 1e-12
        e = eps.copy(); d = O.ragged_eps_update(stats, ones, twos, miss, e, b0, b1, 0.0)
        assert np.max(np.abs((e - eps) - delta)) <= 1e-12 and abs(d - es) <= 1e-12


def test_chain_matches_reference_with_missing_and_groups():
    O.build()
    rng = np.random.default_rng(5)
    N, M = 501, 120
    bed = _random_bed(rng, N, M, 0.02)
    y = rng.normal(size=N).astype(np.float32)
    groups = (np.arange(M) % 2).astype(np.int32)
    S = np.array([[0.001, 0.01, 0.1], [0.002, 0.02, 0.2]])
    for layout, name in ((0, "dense"), (1, "ragged")):
        arr, _ = R.run_chain(bed, N, M, y, S, seed=3, iterations=4, layout=layout, groups=groups)
        for it in range(4):
            mdl = O.ModelArrays(N, S, arr["pi_in"][it], arr["sigmaG_in"][it], arr["sigmaE_in"][it])
            eps = arr["eps_in"][it].copy(); beta = arr["beta_in"][it].copy()
            comp = arr["comp_out"][it - 1].copy() if it else np.zeros(M)
            acum = arr["acum_out"][it - 1].copy() if it else np.zeros(M)
            v, bs, es = O.sweep(name, bed, M, arr["order"][it], arr["rand_uz"][it], mdl, groups, eps, beta, comp, acum,
                                eps_sum=arr["epssum_in"][it])
            assert np.max(np.abs(beta - arr["beta_out"][it])) <= 1e-11
            assert (comp == arr["comp_out"][it]).all() and (v == arr["v_out"][it]).all()
            assert np.max(np.abs(acum - arr["acum_out"][it])) <= 1e-11
            assert np.max(np.abs(eps - arr["eps_out"][it])) <= 1e-10
            assert np.max(np.abs(bs - arr["betasq_out"][it])) <= 1e-11
            assert abs(es - arr["epssum_out"][it]) <= 1e-9


@pytest.mark.parametrize("seed", range(8))
def test_chain_matches_reference_on_random_configurations(seed):
    """Random shapes of the model -- individuals (also not a multiple of 4), markers, groups, mixture components, missing rate,
    all three layouts: every sweep of the reference's chain is replayed by the restatement from the reference's inputs."""
    O.build()
    rng = np.random.default_rng(100 + seed)
    N = int(rng.choice([33, 130, 402, 777])); M = int(rng.integers(20, 90))
    nG = int(rng.integers(1, 4)); K1 = int(rng.integers(1, 5))
    p_miss = float(rng.choice([0.0, 0.01, 0.15]))
    bed = _random_bed(rng, N, M, p_miss)
    beta_true = np.where(rng.random(M) < 0.2, rng.normal(size=M), 0.0)
    cols = np.array([np.nan_to_num(O.dense_column(bed[j], N)) for j in range(M)])
    y = (beta_true @ cols * 0.3 + rng.normal(size=N)).astype(np.float32)
    groups = rng.integers(0, nG, M).astype(np.int32)
    groups[:nG] = np.arange(nG)                                          # every group has a marker
    S = np.sort(rng.uniform(1e-4, 0.5, size=(nG, K1)), axis=1)
    iters = 3
    for layout, name in ((0, "dense"), (1, "ragged"), (2, "eigen")):
        arr, _ = R.run_chain(bed, N, M, y, S, seed=seed + 1, iterations=iters, layout=layout, groups=groups if nG > 1 else None)
        for it in range(iters):
            mdl = O.ModelArrays(N, S, arr["pi_in"][it], arr["sigmaG_in"][it], arr["sigmaE_in"][it])
            eps = arr["eps_in"][it].copy(); beta = arr["beta_in"][it].copy()
            comp = arr["comp_out"][it - 1].copy() if it else np.zeros(M)
            acum = arr["acum_out"][it - 1].copy() if it else np.zeros(M)
            v, bs, es = O.sweep(name, bed, M, arr["order"][it], arr["rand_uz"][it], mdl, groups if nG > 1 else np.zeros(M, np.int32), eps, beta, comp, acum,
                                eps_sum=arr["epssum_in"][it])
            scale = 1.0 + np.abs(arr["beta_out"][it]).max()
            assert np.max(np.abs(beta - arr["beta_out"][it])) <= 1e-10 * scale, (name, it)
            assert (comp == arr["comp_out"][it]).all() and (v == arr["v_out"][it]).all(), (name, it)
            assert np.max(np.abs(acum - arr["acum_out"][it])) <= 1e-10
            assert np.max(np.abs(eps - arr["eps_out"][it])) <= 1e-9 * (1.0 + np.abs(arr["eps_out"][it]).max())

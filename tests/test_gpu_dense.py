"""GPU parity tests, dense 2-bit layout, through the C ABI (include/brr_gpu.h).

Tier 1: computeNum / epsilon change per marker vs the reference's golden numbers (<= 1e-10 rel).
Tier 2: 10-iteration replay of the reference chain consuming the recorded RNG table (<= 1e-8).
Plus: CUDA path vs the C oracle on seeded random inputs, edge cases, async mode, block sizes.
"""
import os

import numpy as np
import pytest

import oracle_lib as O
from bayesrrcmd_b200 import Engine
from helpers import Golden, check_iteration, num_tol, replay

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["v4", "v5", "v6"])
def dense_kernel(request, monkeypatch):
    """Every test of this module runs against both dense sweep kernels (BRR_GPU_DENSE_KERNEL is read at brr_gpu_create): version 4
    (on-demand Gram rows; also the asynchronous and the sharded path), version 5 (steady-state pipeline; synchronous, one GPU) and
    version 6 (the same with the walk distributed over the CTAs)."""
    if request.param in ("v5", "v6"):
        cs = getattr(request.node, "callspec", None)
        prm = cs.params if cs else {}
        sharded_ok = request.param == "v6"         # version 6 also runs sharded over several GPUs (synchronous mode)
        if prm.get("tokens", 0) or (prm.get("world", 1) > 1 and not sharded_ok) or request.node.name.startswith(("test_world_gt_1", "test_mu_update", "test_synthetic", "test_tier1")):
            pytest.skip("asynchronous / sharded / marker-level paths do not run the version-5 / 6 kernels")
    monkeypatch.setenv("BRR_GPU_DENSE_KERNEL", request.param)
    return request.param


def make_engine(g, layout="dense", **kw):
    e = Engine(g.N, g.M, g.K, g.nG, layout=layout, **kw)
    e.upload_bed(g.bed)
    e.finalize()
    e.set_model(g.S, g.groups)
    return e


@pytest.mark.parametrize("name", ["uk10k", "gauss", "missing_small"])
def test_tier1_compute_num_and_delta(name):
    g = Golden(name)
    t = g.tier1()
    e = make_engine(g)
    lut = e.dense_lut()
    eps = t["eps"]; en = np.linalg.norm(eps)
    e.set_state(eps=eps)
    for n, j in enumerate(t["cols"]):
        for c in range(4):
            if not np.isnan(t["lut"][n][c]):
                assert abs(lut[j][c] - t["lut"][n][c]) <= 1e-12 * (1 + abs(lut[j][c])), (j, c)
        b0, b1 = float(t["beta_old"][n]), float(t["beta_new"][n])
        num = e.compute_num(int(j), b0)
        assert abs(num - t["num_dense"][n]) <= num_tol(t["num_dense"][n], np.sqrt(g.N - 1.0), en), (j, num, t["num_dense"][n])
    for n in range(len(t["delta_dense"])):
        j = int(t["cols"][n])
        e.set_state(eps=eps)
        e.apply_delta(j, float(t["beta_old"][n]), float(t["beta_new"][n]))
        got = e.get_state(beta=False, comp=False, acum=False)[0] - eps
        assert np.max(np.abs(got - t["delta_dense"][n])) <= 1e-12


def _replay_gpu(g, e, layout, iters, async_tokens=0):
    chain = g.chain(layout)
    e.set_state(eps=np.zeros(g.N), beta=np.zeros(g.M), comp=np.zeros(g.M), acum=np.zeros(g.M))
    outs = []

    def sweep(it, order, rand_uz, pi, sigmaG, sigmaE, eps_sum):
        v, bs, out = e.sweep(order, rand_uz, pi, sigmaG, sigmaE, eps_sum, async_tokens)
        outs.append(out)
        return v, bs, out.eps_sum

    return chain, replay(chain, iters, sweep, e.get_state, lambda x: e.set_state(eps=x)), outs


@pytest.mark.parametrize("name,block", [("uk10k", 16), ("uk10k", 32), ("groups", 16), ("gauss", 8), ("missing_small", 16),
                                        ("missing_small", 5), ("missing_small", 1)])
def test_tier2_replay_reference_chain(name, block):
    g = Golden(name)
    e = make_engine(g, block_markers=block)
    chain, it_res, outs = _replay_gpu(g, e, "dense", len(g.chain("dense")["mu"]))
    for it, res in it_res:
        check_iteration(chain, it, res, tol=1e-8)
        assert outs[-1].n_nonskipped == chain["nonskipped"][it]
        assert abs(outs[-1].eps_sqnorm - chain["chk_eps2"][it]) <= 1e-8 * chain["chk_eps2"][it]
        assert abs(outs[-1].eps_sum - chain["chk_eps"][it]) <= 1e-8 * (1 + abs(chain["chk_eps"][it]))


def _random_problem(seed, N, M, p_miss, nG=1, K=4):
    rng = np.random.default_rng(seed)
    maf = rng.uniform(0.02, 0.5, M)
    geno = rng.binomial(2, maf[:, None], size=(M, N))
    code = np.where(rng.random((M, N)) < p_miss, 1, np.array([3, 2, 0])[geno]).astype(np.uint8)
    bed = np.zeros((M, (N + 3) // 4), dtype=np.uint8)
    for k in range(4):
        sub = code[:, k::4]
        bed[:, :sub.shape[1]] |= (sub << (2 * k)).astype(np.uint8)
    eps = rng.normal(size=N)
    S = np.sort(rng.uniform(1e-4, 1e-1, size=(nG, K - 1)), axis=1)
    groups = rng.integers(0, nG, M).astype(np.int32) if nG > 1 else None
    pi = rng.dirichlet(np.r_[20.0, np.ones(K - 1)], size=nG)
    sigmaG = rng.uniform(0.1, 0.9, nG)
    return dict(rng=rng, bed=bed, eps=eps, S=S, groups=groups, pi=pi, sigmaG=sigmaG, sigmaE=0.6, N=N, M=M, K=K, nG=nG)


@pytest.mark.parametrize("N,M,p_miss,nG,K,block,tokens", [
    (64, 40, 0.0, 1, 4, 16, 0),        # smallest slice
    (1000, 300, 0.01, 2, 3, 16, 0),
    (5003, 200, 0.02, 1, 5, 32, 0),    # ragged N, several CTAs
    (70001, 96, 0.001, 1, 4, 16, 0),   # every SM busy
    (3000, 257, 0.01, 2, 4, 16, 20),   # asynchronous, 20 tokens (reference default)
    (3000, 100, 0.0, 1, 2, 4, 0),      # K = 2
    (2500, 640, 0.3, 3, 4, 64, 0),     # 30% missing: missing lists far beyond the 15 kept in a row's tail (rescans, long-list pairs)
    (4000, 700, 0.05, 1, 8, 64, 0),    # K = 8 (general sampler), full blocks, carry across blocks
    (9000, 520, 0.002, 5, 4, 48, 0),   # blocks of 48 (3 marker tiles), 5 groups
    (3000, 400, 0.01, 2, 4, 64, 64),   # asynchronous with 64 tokens: carry only
    (1200, 300, 0.6, 1, 3, 64, 0),     # mostly missing genotypes
])
def test_sweep_matches_oracle_on_random_inputs(N, M, p_miss, nG, K, block, tokens):
    _compare_sweeps_with_oracle(N, M, p_miss, nG, K, block, tokens)


@pytest.mark.parametrize("N,max_ctas,slice_inds", [
    (3000, 10, 320),      # 5 k-steps per slice: two word pairs + a half pair, dot warps without any item
    (3000, 7, 448),       # 7 k-steps: 3 pairs + half
    (8000, 10, 832),      # 13 k-steps: the half pair lands on another warp of the tile
    (11000, 7, 1600),     # 25 k-steps: several items per dot warp (the software-pipelined loop)
    (9000, 5, 1856),      # 29 k-steps
])
def test_slice_shapes_of_the_mma_loops(N, max_ctas, slice_inds):
    """The dot and Gram-row loops split a slice's word pairs over the warps of a marker tile (4+4+3+3) and handle an odd
    number of 64-individual k-steps by a half pair: every split is compared with the oracle (the full-size tests cover 53)."""
    info = _compare_sweeps_with_oracle(N, 200, 0.02, 2, 4, 64, 0, max_ctas=max_ctas)
    assert info["slice_inds"] == slice_inds


def _compare_sweeps_with_oracle(N, M, p_miss, nG, K, block, tokens, max_ctas=0):
    O.build()
    pr = _random_problem(N * 7 + M, N, M, p_miss, nG, K)
    rng = pr["rng"]
    e = Engine(N, M, K, nG, block_markers=block, max_ctas=max_ctas)
    e.upload_bed(pr["bed"]); e.finalize(); e.set_model(pr["S"], pr["groups"])
    beta = np.where(rng.random(M) < 0.3, rng.normal(size=M) * 0.05, 0.0)
    comp = (beta != 0).astype(np.float64); acum = rng.random(M)
    eps_o = pr["eps"].copy(); beta_o = beta.copy(); comp_o = comp.copy(); acum_o = acum.copy()
    e.set_state(eps=pr["eps"], beta=beta, comp=comp, acum=acum)
    for sweep_i in range(2):
        order = rng.permutation(M).astype(np.uint32)
        if sweep_i == 1:
            order = order[: M - 7]          # a partial sweep is legal at this boundary
        rand_uz = np.c_[rng.random(M), rng.normal(size=M)]
        mdl = O.ModelArrays(N, pr["S"], pr["pi"], pr["sigmaG"], pr["sigmaE"])
        v_o, bs_o, _ = O.sweep("dense", pr["bed"], M, order, rand_uz, mdl, pr["groups"], eps_o, beta_o, comp_o, acum_o,
                               tokens=tokens)
        v, bs, out = e.sweep(order, rand_uz, pr["pi"], pr["sigmaG"], pr["sigmaE"], 0.0, tokens)
        eps, b, c, a = e.get_state()
        assert (c == comp_o).all()
        assert (v == v_o).all()
        assert np.max(np.abs(b - beta_o)) <= 1e-9
        assert np.max(np.abs(a - acum_o)) <= 1e-9
        assert np.max(np.abs(eps - eps_o)) <= 1e-9
        assert np.max(np.abs(bs - bs_o)) <= 1e-9 * (1 + np.abs(bs_o).max())
        assert abs(out.eps_sqnorm - (eps_o ** 2).sum()) <= 1e-9 * (eps_o ** 2).sum()
    return e.info()


def test_empty_and_single_marker_sweeps():
    pr = _random_problem(11, 777, 9, 0.05)
    e = Engine(777, 9, 4, 1)
    e.upload_bed(pr["bed"]); e.finalize(); e.set_model(pr["S"])
    e.set_state(eps=pr["eps"])
    rand_uz = np.c_[np.full(9, 0.5), np.zeros(9)]
    v, bs, out = e.sweep(np.zeros(0, np.uint32), rand_uz, pr["pi"], pr["sigmaG"], 0.5)
    assert v.sum() == 0 and out.n_blocks == 0
    assert np.array_equal(e.get_state()[0], pr["eps"])
    v, bs, out = e.sweep(np.array([3], np.uint32), rand_uz, pr["pi"], pr["sigmaG"], 0.5)
    assert v.sum() == 1 and out.n_blocks == 1


def test_mu_update_primitives():
    pr = _random_problem(3, 4099, 4, 0.0)
    e = Engine(4099, 4, 4, 1)
    e.set_state(eps=pr["eps"])
    s = e.eps_add_scalar(0.25)
    ref = pr["eps"] + 0.25
    assert abs(s - ref.sum()) <= 1e-10 * (1 + abs(ref.sum()))
    assert np.array_equal(e.get_state()[0], ref)
    s2, q2 = e.eps_reduce()
    assert abs(q2 - (ref ** 2).sum()) <= 1e-10 * (ref ** 2).sum()


def test_errors_are_reported_not_thrown():
    from bayesrrcmd_b200 import BrrGpuError
    pr = _random_problem(5, 300, 20, 0.0)
    e = Engine(300, 20, 4, 1)
    rand_uz = np.zeros((20, 2))
    with pytest.raises(BrrGpuError, match="upload"):
        e.sweep(np.arange(20, dtype=np.uint32), rand_uz, pr["pi"], pr["sigmaG"], 0.5)
    e.upload_bed(pr["bed"]); e.finalize(); e.set_model(pr["S"])
    with pytest.raises(BrrGpuError, match="twice"):
        e.sweep(np.array([1, 1], np.uint32), rand_uz, pr["pi"], pr["sigmaG"], 0.5)
    with pytest.raises(BrrGpuError, match="out of range"):
        e.sweep(np.array([20], np.uint32), rand_uz, pr["pi"], pr["sigmaG"], 0.5)


def test_synthetic_generator_matches_its_host_twin():
    from bayesrrcmd_b200.capi import synth_bed_host
    N, M = 5000, 64
    e = Engine(N, M, 4, 1)
    e.synth_dense(seed=42, p_missing=0.01)
    e.finalize()
    bed = synth_bed_host(42, N, 0, M, p_missing=0.01)
    e2 = Engine(N, M, 4, 1)
    e2.upload_bed(bed); e2.finalize()
    assert (e.counts() == e2.counts()).all()
    assert e.counts()[:, 3].sum() > 0
    eps = np.random.default_rng(0).normal(size=N)
    e.set_state(eps=eps); e2.set_state(eps=eps)
    for j in (0, 17, 63):
        assert e.compute_num(j, 0.1) == e2.compute_num(j, 0.1)


def _shard_bounds(N, world):
    per = ((N + world - 1) // world + 63) // 64 * 64
    return [(min(N, r * per), min(N, (r + 1) * per)) for r in range(world)]


def _shard_bed(bed, lo, hi):
    """Rows lo:hi of a PLINK payload (lo is a multiple of 4)."""
    assert lo % 4 == 0
    return np.ascontiguousarray(bed[:, lo // 4:(hi + 3) // 4])


@pytest.mark.parametrize("world,N,M,p_miss,block,tokens", [(2, 3000, 150, 0.01, 64, 0), (3, 5000, 100, 0.0, 16, 0), (2, 4000, 300, 0.02, 64, 20)])
def test_sharded_individuals_exchange_on_one_device(world, N, M, p_miss, block, tokens):
    """The in-kernel cross-shard exchange (the multi-GPU path) with all shards on ONE device: the kernels of
    the shards run concurrently on their own streams and exchange through 'peer' pointers that happen to be
    local.  The chain must be BIT-identical to the unsharded one (the cross-shard sums are integers)."""
    _sharded_exchange(world, N, M, p_miss, block, tokens, devices=[0] * world)


@pytest.mark.parametrize("world,N,M,p_miss,block,tokens", [(2, 200000, 400, 0.002, 64, 0), (2, 60000, 300, 0.01, 64, 20), (4, 300000, 300, 0.001, 64, 0)])
def test_sharded_individuals_exchange_on_several_devices(world, N, M, p_miss, block, tokens):
    """The same bit-identity with every shard on its OWN GPU (NVLink peer memory between real devices).  Needs `world` visible
    devices: skipped on a single-GPU box, run by `gpurun --gpus N`."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    _sharded_exchange(world, N, M, p_miss, block, tokens, devices=list(range(world)))


def _sharded_exchange(world, N, M, p_miss, block, tokens, devices):
    # version 6 runs deeper across GPUs (7 blocks) than on one (4); chains of different depths agree to rounding only, so the unsharded
    # comparison run is given the sharded depth: at equal depth the sharded chain is BIT-identical (integer cross-shard sums)
    os.environ["BRR_GPU_V6_DEPTH"] = "7"
    try:
        _sharded_exchange_at_depth(world, N, M, p_miss, block, tokens, devices)
    finally:
        del os.environ["BRR_GPU_V6_DEPTH"]


def _sharded_exchange_at_depth(world, N, M, p_miss, block, tokens, devices):
    O.build()
    pr = _random_problem(N + M + world, N, M, p_miss)
    rng = pr["rng"]
    ref = Engine(N, M, pr["K"], 1, block_markers=block)
    ref.upload_bed(pr["bed"]); ref.finalize(); ref.set_model(pr["S"])
    bounds = _shard_bounds(N, world)
    shards = []
    for r, (lo, hi) in enumerate(bounds):
        e = Engine(hi - lo, M, pr["K"], 1, block_markers=block, n_inds_total=N, rank=r, world=world, device=devices[r])
        e.upload_bed(_shard_bed(pr["bed"], lo, hi))
        shards.append(e)
    counts = sum(e.counts().astype(np.int64) for e in shards).astype(np.uint32)
    assert (counts == ref.counts()).all()
    for e in shards:
        e.finalize(counts); e.set_model(pr["S"])
    Engine.attach_local(shards)
    beta = np.where(rng.random(M) < 0.3, rng.normal(size=M) * 0.05, 0.0)
    ref.set_state(eps=pr["eps"], beta=beta, comp=(beta != 0).astype(float), acum=np.zeros(M))
    for e, (lo, hi) in zip(shards, bounds):
        e.set_state(eps=pr["eps"][lo:hi], beta=beta, comp=(beta != 0).astype(float), acum=np.zeros(M))
    for sweep_i in range(3):
        order = rng.permutation(M).astype(np.uint32)
        rand_uz = np.c_[rng.random(M), rng.normal(size=M)]
        v0, bs0, out0 = ref.sweep(order, rand_uz, pr["pi"], pr["sigmaG"], pr["sigmaE"], 0.0, tokens)
        for e in shards:
            e.sweep_stage(order, rand_uz, pr["pi"], pr["sigmaG"], pr["sigmaE"], 0.0, tokens)
        for e in shards:
            e.sweep_run()                      # asynchronous: all shards must be in flight together
        res = [e.sweep_fetch() for e in shards]
        eps0, b0, c0, a0 = ref.get_state()
        eps_sh = np.concatenate([e.get_state()[0] for e in shards])
        assert np.array_equal(eps_sh, eps0)
        for e, (v, bs, out) in zip(shards, res):
            _, b, c, a = e.get_state()
            assert np.array_equal(b, b0) and np.array_equal(c, c0) and np.array_equal(a, a0)
            assert np.array_equal(v, v0)
            if os.environ.get("BRR_GPU_DENSE_KERNEL") == "v6":      # sum(beta^2) is added up per owner CTA: the order depends on the number of CTAs
                assert np.max(np.abs(bs - bs0)) <= 1e-13 * (1 + np.abs(bs0).max())
            else:
                assert np.array_equal(bs, bs0)
            assert out.n_nonskipped == out0.n_nonskipped
        assert abs(sum(o.eps_sqnorm for _, _, o in res) - out0.eps_sqnorm) <= 1e-9 * out0.eps_sqnorm


def test_world_gt_1_requires_attach():
    from bayesrrcmd_b200 import BrrGpuError
    pr = _random_problem(5, 640, 20, 0.0)
    e = Engine(320, 20, 4, 1, n_inds_total=640, rank=0, world=2)
    e.upload_bed(_shard_bed(pr["bed"], 0, 320)); e.finalize(); e.set_model(pr["S"])
    with pytest.raises(BrrGpuError, match="attach"):
        e.sweep(np.arange(20, dtype=np.uint32), np.zeros((20, 2)), pr["pi"], pr["sigmaG"], 0.5)


def test_full_size_properties_n500k():
    """At the bench's individual count (N = 500 000, one GPU; M = 1024 keeps it to seconds) the oracle is too slow to replay, so
    the sweep is checked through size-independent properties: (1) the residual identity eps_end = eps_start + sum_j (beta_old_j -
    beta_new_j) x_j, rebuilt marker by marker with the parity hook brr_gpu_apply_delta; (2) the sweep's sum(eps) and ||eps||^2
    against the fetched residual; (3) the counts: v sums to the markers visited, n_nonskipped matches the beta vectors; (4) the
    chain is bit-identical when repeated from the same inputs, and for another block size (the integer reductions do not
    depend on the blocking)."""
    N, M, K = 500000, 1024, 4
    rng = np.random.default_rng(123)
    S = np.array([[0.01, 0.001, 0.0001]])
    y = rng.standard_normal(N); y = (y - y.mean()) / y.std(ddof=1)
    pi = np.array([[0.7, 0.1, 0.1, 0.1]]); sigmaG = np.array([0.3]); sigmaE = 0.7
    sweeps = [(rng.permutation(M).astype(np.uint32), np.c_[rng.random(M), rng.standard_normal(M)]) for _ in range(2)]

    def run(block):
        e = Engine(N, M, K, 1, block_markers=block)
        e.synth_dense(7, 0.01, 0.5, False, 0.001)
        e.finalize(); e.set_model(S)
        e.set_state(eps=y, beta=np.zeros(M), comp=np.zeros(M), acum=np.zeros(M))
        hist = []
        for order, uz in sweeps:
            b_old = e.get_state(eps=False)[1].copy()
            v, bs, out = e.sweep(order, uz, pi, sigmaG, sigmaE)
            eps, b_new, comp, acum = e.get_state()
            hist.append(dict(b_old=b_old, b_new=b_new.copy(), eps=eps.copy(), comp=comp.copy(), acum=acum.copy(), v=v.copy(), bs=bs.copy(), out=out))
        return e, hist

    e, hist = run(0)
    eps_ref = y.copy()
    for h in hist:
        out = h["out"]
        assert h["v"].sum() == M
        nonskipped = int(((h["b_old"] != 0) | (h["b_new"] != 0)).sum())
        assert out.n_nonskipped == nonskipped and nonskipped > 0
        assert (h["v"][0, 1:].sum() == (h["b_new"] != 0).sum())
        assert abs(out.eps_sqnorm - (h["eps"] ** 2).sum()) <= 1e-11 * (h["eps"] ** 2).sum()
        assert abs(out.eps_sum - h["eps"].sum()) <= 1e-9 * (1 + abs(h["eps"].sum()))
        assert abs(h["bs"][0] - (h["b_new"] ** 2).sum()) <= 1e-12 * (1 + (h["b_new"] ** 2).sum())
        # (1) rebuild the residual with the marker-level hook on the same store
        e.set_state(eps=eps_ref)
        for j in np.nonzero(h["b_old"] != h["b_new"])[0]:
            e.apply_delta(int(j), float(h["b_old"][j]), float(h["b_new"][j]))
        eps_ref = e.get_state(beta=False, comp=False, acum=False)[0]
        assert np.max(np.abs(eps_ref - h["eps"])) <= 1e-9
        e.set_state(eps=h["eps"])
    _, hist2 = run(0)
    _, hist3 = run(32)
    for h, h2, h3 in zip(hist, hist2, hist3):
        for k in ("b_new", "eps", "comp", "acum", "v", "bs"):
            assert np.array_equal(h[k], h2[k]), k                       # (4) repeatable bit for bit
        assert np.array_equal(h["comp"], h3["comp"]) and np.array_equal(h["v"], h3["v"])
        assert np.max(np.abs(h["b_new"] - h3["b_new"])) <= 1e-10 and np.max(np.abs(h["eps"] - h3["eps"])) <= 1e-9


def test_replay_reference_chain_at_n500k():
    """Tier 2 at the bench's individual count: the REFERENCE ITSELF (oracle/_ref, built from the reference's sources) runs two
    Gibbs sweeps over N = 500 000 individuals x 96 markers of the bench's synthetic data set; the GPU sweep replays each of them
    from the recorded inputs (residual, betas, order, (u, z), pi, sigmaG, sigmaE) and must land on the reference's betas,
    components, acum and residual within 1e-8."""
    import ref_lib as R
    from bayesrrcmd_b200.capi import synth_bed_host
    if not R.ref_available():
        pytest.skip("oracle/_ref is not built")
    N, M = 500000, 96
    S = np.array([[0.01, 0.001, 0.0001]])
    bed = synth_bed_host(20260101, N, 0, M, p_missing=0.001)
    rng = np.random.default_rng(5)
    y = rng.standard_normal(N).astype(np.float32)
    arr, _ = R.run_chain(bed, N, M, y, S, seed=3, iterations=2, layout=0)
    e = Engine(N, M, 4, 1)
    e.upload_bed(bed); e.finalize(); e.set_model(S)
    comp_in = np.zeros(M)
    for it in range(2):
        e.set_state(eps=arr["eps_in"][it], beta=arr["beta_in"][it], comp=comp_in, acum=np.zeros(M))
        v, bs, out = e.sweep(arr["order"][it].astype(np.uint32), arr["rand_uz"][it], arr["pi_in"][it], arr["sigmaG_in"][it],
                             float(arr["sigmaE_in"][it]))
        eps, beta, comp, acum = e.get_state()
        assert np.array_equal(comp, arr["comp_out"][it])
        assert np.array_equal(v, arr["v_out"][it])
        for got, ref, what in ((beta, arr["beta_out"][it], "beta"), (acum, arr["acum_out"][it], "acum"), (eps, arr["eps_out"][it], "eps"),
                               (bs, arr["betasq_out"][it], "betasqnG")):
            err = np.max(np.abs(got - ref) / (1.0 + np.abs(ref)))
            assert err <= 1e-8, "%s differs at iteration %d: %.3e" % (what, it, err)
        assert (arr["beta_out"][it] != 0).sum() > 0
        comp_in = arr["comp_out"][it]

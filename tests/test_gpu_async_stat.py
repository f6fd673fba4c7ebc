"""GPU: the asynchronous mode (brr_gpu_sweep with async_tokens = 20 = the reference's default analysisTokens) against the
REFERENCE'S OWN asyncppbayes chain -- BayesRBase::processColumnAsync / doThreadSafeUpdates / updateGlobal under the reference's
ParallelGraph (TBB), run by oracle/make_golden_async.py into tests/golden/async_uk10k.npz.  The asynchronous chain is
scheduling-dependent by construction, so the comparison is statistical: posterior means of sigmaE, sigmaG and sum(beta^2) (the number of
non-zero effects mixes too slowly to be a test at this length) over iterations 100..299 of a 300-iteration chain on the uk10k fixture, within Monte-Carlo error (batch
means).  NOTE what this can and cannot show: under TBB the reference applies a marker's residual update as soon as that marker is
done (serial global-update node, src/parallelgraph.cpp:92-105), so its effective staleness is at most the number of worker threads;
the GPU mode lets all T markers of a round read the same residual (the upper bound of that staleness).  At this chain length the
posterior summaries of the synchronous chain, the reference's asynchronous chain and the GPU's agree within Monte-Carlo error."""
import numpy as np
import pytest

from bayesrrcmd_b200 import Engine
from helpers import GOLDEN, Golden, run_host_chain

pytestmark = pytest.mark.gpu

BURN = 100


def _gpu_chain(g, iters, tokens, seed):
    e = Engine(g.N, g.M, g.K, 1, layout="dense")
    e.upload_bed(g.bed); e.finalize(); e.set_model(g.S)
    y = g.y.astype(np.float32).astype(np.float64)
    y = y - y.mean(); y = y / np.sqrt((y ** 2).sum() / (g.N - 1))
    e.set_state(eps=y, beta=np.zeros(g.M), comp=np.zeros(g.M), acum=np.zeros(g.M))

    def sweep(order, rand_uz, pi, sigmaG, sigmaE):
        v, bs, out = e.sweep(order, rand_uz, pi, sigmaG, sigmaE, 0.0, tokens)
        beta = e.get_state(eps=False, comp=False, acum=False)[1]
        return v[0], float(bs[0]), out.eps_sqnorm, float((beta != 0).sum())

    return run_host_chain(g.N, g.M, g.S, y, iters, seed, sweep, e.eps_add_scalar)


def _mc_err(x):
    """Monte-Carlo standard error of the mean of a correlated series (batch means, 10 batches)."""
    b = x[: len(x) // 10 * 10].reshape(10, -1).mean(axis=1)
    return b.std(ddof=1) / np.sqrt(10)


def test_async_posterior_means_match_the_reference_parallelgraph_chain():
    z = np.load(GOLDEN + "/async_uk10k.npz")
    g = Golden("uk10k")
    iters = int(z["iters"])
    gpu = _gpu_chain(g, iters, int(z["tokens"]), seed=11)
    gpu_sync = _gpu_chain(g, iters, 0, seed=12)
    for k, rel in (("sigmaE", 0.02), ("sigmaG", 0.06), ("betasq", 0.06)):
        ref = z["async__" + k][BURN:]; ref_sync = z["sync__" + k][BURN:]
        got = gpu[k][BURN:]; got_sync = gpu_sync[k][BURN:]
        tol = 4.0 * np.hypot(_mc_err(ref), _mc_err(got)) + rel * abs(ref.mean())
        assert abs(got.mean() - ref.mean()) <= tol, "%s: GPU async %.4f vs reference async %.4f (tol %.4f)" % (k, got.mean(), ref.mean(), tol)
        tol_s = 4.0 * np.hypot(_mc_err(ref_sync), _mc_err(got_sync)) + rel * abs(ref_sync.mean())
        assert abs(got_sync.mean() - ref_sync.mean()) <= tol_s, "%s: GPU sync %.4f vs reference sync %.4f" % (k, got_sync.mean(), ref_sync.mean())

"""CPU: the plain-C restatement's asynchronous mode (oracle/bayesr_oracle.c, bounded staleness of T markers per synchronisation)
against the reference's own ParallelGraph chain (tests/golden/async_uk10k.npz, oracle/make_golden_async.py): posterior means over a
shorter chain, within Monte-Carlo error.  Pins the model of asynchrony the GPU tests compare with bit for bit."""
import numpy as np

import oracle_lib as O
from helpers import GOLDEN, Golden, run_host_chain


def test_oracle_async_matches_reference_parallelgraph_statistics():
    O.build()
    z = np.load(GOLDEN + "/async_uk10k.npz")
    g = Golden("uk10k")
    iters, burn = 120, 50
    y = g.y.astype(np.float32).astype(np.float64)
    y = y - y.mean(); y = y / np.sqrt((y ** 2).sum() / (g.N - 1))
    eps = y.copy(); beta = np.zeros(g.M); comp = np.zeros(g.M); acum = np.zeros(g.M)

    def add(c):
        eps[:] = eps + c
        return float(eps.sum())

    def sweep(order, rand_uz, pi, sigmaG, sigmaE):
        mdl = O.ModelArrays(g.N, g.S, pi, sigmaG, sigmaE)
        v, bs, _ = O.sweep("dense", g.bed, g.M, order, rand_uz, mdl, None, eps, beta, comp, acum, tokens=int(z["tokens"]))
        return v[0], float(bs[0]), float((eps ** 2).sum()), float((beta != 0).sum())

    got = run_host_chain(g.N, g.M, g.S, y, iters, 5, sweep, add)
    def mc(x):      # Monte-Carlo standard error of the mean of a correlated series (batch means)
        b = x[: len(x) // 10 * 10].reshape(10, -1).mean(axis=1)
        return b.std(ddof=1) / np.sqrt(10)

    for k, rel in (("sigmaE", 0.02), ("sigmaG", 0.06), ("betasq", 0.06)):
        ref = z["async__" + k][100:]
        tol = 4.0 * np.hypot(mc(ref), mc(got[k][burn:])) + rel * abs(ref.mean())
        assert abs(got[k][burn:].mean() - ref.mean()) <= tol, (k, got[k][burn:].mean(), ref.mean(), tol)

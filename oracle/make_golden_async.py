"""TEST INFRASTRUCTURE ONLY: the statistical pin of the asynchronous mode.

Runs the UNMODIFIED reference's asyncppbayes path -- BayesRBase::processColumnAsync / doThreadSafeUpdates / updateGlobal
(src/BayesRBase.cpp:394-523) under the reference's own ParallelGraph (src/parallelgraph.cpp, TBB flow graph, 20 analysis tokens,
all host cores; built by oracle/Makefile from /root/reference) -- for a long chain on the uk10k fixture, and the synchronous
chain from the same seed beside it, and keeps the per-iteration scalars whose posterior means the GPU's asynchronous mode must
reproduce within Monte-Carlo error (tests/test_gpu_async_stat.py): sigmaE, sigmaG, sum(beta^2), number of non-zero effects, mu.

    python oracle/make_golden_async.py        ->  tests/golden/async_uk10k.npz   (a few KB)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
import ref_lib as R          # noqa: E402
from helpers import Golden   # noqa: E402

ITERS, TOKENS, SEED = 300, 20, 1


def scalars(arr):
    return dict(mu=arr["mu"], sigmaE=arr["sigmaE_post"], sigmaG=arr["sigmaG_post"][:, 0], betasq=arr["betasq_out"][:, 0],
                nonzero=(arr["beta_out"] != 0).sum(axis=1).astype(np.float64), v=arr["v_out"][:, 0, :])


def main():
    g = Golden("uk10k")
    out = dict(iters=ITERS, tokens=TOKENS, seed=SEED)
    for name, graph in (("async", "parallel"), ("sync", "sequential")):
        arr, tm = R.run_chain(g.bed, g.N, g.M, g.y.astype(np.float32), g.S, seed=SEED, iterations=ITERS, layout=0, graph=graph,
                              tokens=TOKENS)
        for k, v in scalars(arr).items():
            out["%s__%s" % (name, k)] = v
        print(name, "%.1f s" % tm["total_s"], "mean sigmaE %.4f sigmaG %.4f nonzero %.1f" % (
            out[name + "__sigmaE"][100:].mean(), out[name + "__sigmaG"][100:].mean(), out[name + "__nonzero"][100:].mean()))
    np.savez_compressed(os.path.join(os.path.dirname(HERE), "tests", "golden", "async_uk10k.npz"), **out)


if __name__ == "__main__":
    main()

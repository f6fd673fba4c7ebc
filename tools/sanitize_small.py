"""A small dense chain for compute-sanitizer (memcheck / racecheck): a few sweeps with updates, missing genotypes, groups,
partial last block, synchronous and asynchronous.  usage: compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayesrrcmd_b200 import Engine

rng = np.random.default_rng(3)
N, M, K, nG = 1500, 200, 4, 2
maf = rng.uniform(0.05, 0.5, M)
geno = rng.binomial(2, maf[:, None], size=(M, N))
code = np.where(rng.random((M, N)) < 0.02, 1, np.array([3, 2, 0])[geno]).astype(np.uint8)
bed = np.zeros((M, (N + 3) // 4), dtype=np.uint8)
for k in range(4):
    sub = code[:, k::4]
    bed[:, :sub.shape[1]] |= (sub << (2 * k)).astype(np.uint8)
S = np.array([[0.01, 0.001, 0.0001], [0.02, 0.002, 0.0002]])
groups = rng.integers(0, nG, M).astype(np.int32)
for tokens in (0, 20):
    e = Engine(N, M, K, nG)
    e.upload_bed(bed); e.finalize(); e.set_model(S, groups)
    beta = np.where(rng.random(M) < 0.2, rng.normal(size=M) * 0.05, 0.0)
    e.set_state(eps=rng.normal(size=N), beta=beta, comp=(beta != 0).astype(float), acum=np.zeros(M))
    pi = rng.dirichlet(np.r_[20.0, np.ones(K - 1)], size=nG); sigmaG = np.array([0.4, 0.6])
    for s in range(2):
        order = rng.permutation(M).astype(np.uint32)[: M - 5 * s]
        v, bs, out = e.sweep(order, np.c_[rng.random(M), rng.normal(size=M)], pi, sigmaG, 0.6, 0.0, tokens)
        print("tokens", tokens, "sweep", s, "nonskipped", out.n_nonskipped, "blocks", out.n_blocks, flush=True)
for layout in ("ragged", "eigen"):
    e = Engine(N, M, K, nG, layout=layout)
    e.upload_bed(bed); e.finalize(); e.set_model(S, groups)
    beta = np.where(rng.random(M) < 0.2, rng.normal(size=M) * 0.05, 0.0)
    eps = rng.normal(size=N)
    e.set_state(eps=eps, beta=beta, comp=(beta != 0).astype(float), acum=np.zeros(M))
    pi = rng.dirichlet(np.r_[20.0, np.ones(K - 1)], size=nG); sigmaG = np.array([0.4, 0.6])
    es = float(eps.sum())
    for s in range(2):
        order = rng.permutation(M).astype(np.uint32)
        v, bs, out = e.sweep(order, np.c_[rng.random(M), rng.normal(size=M)], pi, sigmaG, 0.6, es, 0)
        es = out.eps_sum
        print(layout, "sweep", s, "nonskipped", out.n_nonskipped, "blocks", out.n_blocks, flush=True)
print("done")

#!/usr/bin/env python
"""Grouped mixtures and the asynchronous mode at scale (BASELINE.json config 5 shape: --group with the test.groups layout, S = 2 x 2,
asyncppbayes with a bounded number of markers in flight): dense 2-bit synthetic genotypes, N = 500 000 individuals on ONE GPU,
2 groups assigned by marker parity, S = "0.001,0.01;0.001,0.01" (test/data/test.groups/test.cva), K = 3.  For each token count
(0 = synchronous ppbayes; 20 = the reference's default analysisTokens; 64 = one full block) a genuine chain is run and the sweep
kernel timed.  Prints one JSON line per mode.

    python tools/bench_grouped_async.py [--n-inds N] [--n-markers M] [--steps K] [--warmup W] [--tokens 0,20,64]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayesrrcmd_b200 import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--n-inds", type=int, default=500000)
ap.add_argument("--n-markers", type=int, default=200000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--tokens", default="0,20,64")
a = ap.parse_args()
N, M, K, nG = a.n_inds, a.n_markers, 3, 2
SEED = 20260101
S = np.array([[0.001, 0.01], [0.001, 0.01]])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6650.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def inv_scaled_chisq(rng, dof, scale):
    return 1.0 / rng.gamma(0.5 * dof, 1.0 / (0.5 * dof * scale))


t0 = time.time()
e = Engine(N, M, K, nG, layout="dense")
e.synth_dense(SEED, 0.01, 0.5, False, 0.001)
groups = (np.arange(M) & 1).astype(np.int32)
e.finalize(); e.set_model(S, groups)
setup = time.time() - t0
rng0 = np.random.default_rng(SEED)
causal = rng0.choice(M, 1000, replace=False)
y = rng0.standard_normal(N) * np.sqrt(0.5)
e.set_state(eps=y, beta=np.zeros(M), comp=np.zeros(M), acum=np.zeros(M))
for j, b in zip(causal, rng0.standard_normal(1000) * np.sqrt(0.5 / 1000)):
    e.apply_delta(int(j), float(b), 0.0)
y = e.get_state(beta=False, comp=False, acum=False)[0].astype(np.float32).astype(np.float64)
y -= y.mean(); y /= np.sqrt((y ** 2).sum() / (N - 1))

for tokens in [int(t) for t in a.tokens.split(",")]:
    rng = np.random.default_rng(SEED + 1)
    e.set_state(eps=y, beta=np.zeros(M), comp=np.zeros(M), acum=np.zeros(M))
    pi = np.full((nG, K), 0.5 / K); pi[:, 0] = 0.5                     # BayesRBase::init (reference src/BayesRBase.cpp:43-50)
    sigmaG = rng.random(nG); sigmaE = float((y ** 2).sum()) / N * 0.5; mu = 0.0
    kms, e2e, nsk, blocks = [], [], [], []
    for it in range(a.warmup + a.steps):
        s = e.eps_add_scalar(mu)                                      # updateMu (src/DenseBayesRRmz.cpp:36-42)
        mu = rng.normal(s / N, np.sqrt(sigmaE / N))
        e.eps_add_scalar(-mu)
        uz = np.empty((M, 2)); uz[:, 0] = rng.random(M); uz[:, 1] = rng.standard_normal(M)
        order = rng.permutation(M).astype(np.uint32)
        t1 = time.perf_counter()
        v, bs, out = e.sweep(order, uz, pi, sigmaG, sigmaE, 0.0, tokens)
        _, beta, comp, acum = e.get_state(eps=False)
        t2 = time.perf_counter()
        if it >= a.warmup:
            kms.append(out.kernel_ms); e2e.append(t2 - t1); nsk.append(out.n_nonskipped); blocks.append(out.n_blocks)
        sigmaE = inv_scaled_chisq(rng, 1e-4 + N, (out.eps_sqnorm + 1e-8) / (1e-4 + N))          # src/BayesRBase.cpp:238-247
        for g in range(nG):
            m0 = v[g, 1:].sum()
            sigmaG[g] = inv_scaled_chisq(rng, 1e-4 + m0, (bs[g] * m0 + 1e-8) / (1e-4 + m0))
            pi[g] = rng.dirichlet(v[g] + 1.0)
    ksum = sum(kms) / 1e3
    alg = ((N + 3) // 4) * M
    line = {"metric": "marker updates/sec (%s Gibbs sweep, dense 2-bit, 2 groups, N=%d)" % ("asyncppbayes" if tokens else "ppbayes", N),
            "value": M * a.steps / ksum, "unit": "marker updates/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * ksum / a.steps, "higher_is_better": True, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic dense 2-bit N=%d x M=%d, 2 groups by marker parity, S=0.001,0.01;0.001,0.01, K=3" % (N, M),
                       "analysis_tokens": tokens, "markers_per_round": e.info()["block_markers"] if tokens == 0 else tokens,
                       "nonskipped_fraction": float(np.mean(nsk)) / M, "rounds_per_sweep": float(np.mean(blocks)), "setup_s": setup},
            "e2e": {"value": M * a.steps / sum(e2e), "unit": "marker updates/s"},
            "roofline": {"bound": "hbm", "achieved": alg / (ksum / a.steps) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ksum / a.steps) / 1e9 / peak, "kernel": "sweep_dense_kernel", "algorithmic_bytes_per_launch": alg}}
    print(json.dumps(line), flush=True)

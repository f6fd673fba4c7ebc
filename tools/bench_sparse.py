#!/usr/bin/env python
"""Sparse layouts at scale (SURVEY 8d, config S4 shape per marker): ragged / eigen index-stream sweeps on synthetic genotypes with
a log-uniform MAF spectrum on [1e-3, 0.5], N = 500 000 individuals on ONE GPU.  M is smaller than S4's 2M (the store is built on
the device: 2 bytes per non-zero genotype; M = 100 000 is ~14 GB); the metric is per marker.  Prints one JSON line per layout.

    python tools/bench_sparse.py [--n-inds N] [--n-markers M] [--steps K] [--warmup W] [--layouts ragged,eigen]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayesrrcmd_b200 import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--n-inds", type=int, default=500000)
ap.add_argument("--n-markers", type=int, default=100000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--layouts", default="ragged,eigen")
ap.add_argument("--block", type=int, default=0, help="markers per window (0 = engine default 16, max 32)")
a = ap.parse_args()
N, M, K = a.n_inds, a.n_markers, 4
SEED = 20260101
S = np.array([[0.01, 0.001, 0.0001]])
peak = 6650.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def inv_scaled_chisq(rng, dof, scale):
    return 1.0 / rng.gamma(0.5 * dof, 1.0 / (0.5 * dof * scale))


for layout in a.layouts.split(","):
    t0 = time.time()
    e = Engine(N, M, K, 1, layout=layout, block_markers=a.block)
    e.synth_ragged(SEED, 0.001, 0.5, True, 0.001)
    e.finalize(); e.set_model(S)
    setup = time.time() - t0
    counts = e.counts().astype(np.int64)
    nnz = counts[:, 1:].sum(axis=1)
    rng = np.random.default_rng(SEED)
    causal = rng.choice(M, min(1000, M // 10), replace=False)
    y = rng.standard_normal(N) * np.sqrt(0.5)
    e.set_state(eps=y, beta=np.zeros(M), comp=np.zeros(M), acum=np.zeros(M))
    for j, b in zip(causal, rng.standard_normal(len(causal)) * np.sqrt(0.5 / len(causal))):
        e.apply_delta(int(j), float(b), 0.0)
    y = e.get_state(beta=False, comp=False, acum=False)[0].astype(np.float32).astype(np.float64)
    y -= y.mean(); y /= np.sqrt((y ** 2).sum() / (N - 1))
    e.set_state(eps=y)
    pi = np.full((1, K), 0.5 / K); pi[0, 0] = 0.5
    sigmaG = np.array([rng.random()]); sigmaE = float((y ** 2).sum()) / N * 0.5; mu = 0.0
    eps_sum = float(y.sum())
    kms, e2e, nsk, passes = [], [], [], []
    for it in range(a.warmup + a.steps):
        # SparseBayesRRG::updateMu (reference src/SparseBayesRRG.cpp:106-113) with the tracked sum
        e.eps_add_scalar(mu); eps_sum += mu * N
        mu = rng.normal(eps_sum / N, np.sqrt(sigmaE / N))
        e.eps_add_scalar(-mu); eps_sum -= mu * N
        uz = np.empty((M, 2)); uz[:, 0] = rng.random(M); uz[:, 1] = rng.standard_normal(M)
        order = rng.permutation(M).astype(np.uint32)
        t1 = time.perf_counter()
        v, bs, out = e.sweep(order, uz, pi, sigmaG, sigmaE, eps_sum)
        _, beta, comp, acum = e.get_state(eps=False)
        t2 = time.perf_counter()
        eps_sum = out.eps_sum
        if it >= a.warmup:
            kms.append(out.kernel_ms); e2e.append(t2 - t1); nsk.append(out.n_nonskipped); passes.append(out.n_blocks)
        sigmaE = inv_scaled_chisq(rng, 1e-4 + N, (out.eps_sqnorm + 1e-8) / (1e-4 + N))
        m0 = v[0, 1:].sum()
        sigmaG = np.array([inv_scaled_chisq(rng, 1e-4 + m0, (bs[0] * m0 + 1e-8) / (1e-4 + m0))])
        pi = rng.dirichlet(v[0] + 1.0)[None, :]
    ksum = sum(kms) / 1e3
    alg = int((2 * nnz + 32).sum())                     # 16-bit slice-local indices + the marker's statistics
    line = {"metric": "marker updates/sec (ppbayes Gibbs sweep, %s index lists, N=%d)" % (layout, N), "value": M * a.steps / ksum,
            "unit": "marker updates/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * ksum / a.steps,
            "higher_is_better": True, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic %s N=%d x M=%d, MAF log-uniform [0.001,0.5], 0.1%% missing" % (layout, N, M),
                       "mean_nonzero_fraction": float(nnz.mean() / N), "index_bytes": int(e.info()["genotype_bytes"]),
                       "window_markers": e.info()["block_markers"], "nonskipped_fraction": float(np.mean(nsk)) / M, "grid_rounds_per_sweep": float(np.mean(passes)), "setup_s": setup},
            "e2e": {"value": M * a.steps / sum(e2e), "unit": "marker updates/s"},
            "roofline": {"bound": "hbm", "achieved": alg / (ksum / a.steps) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ksum / a.steps) / 1e9 / peak, "kernel": "sweep_ragged_kernel", "algorithmic_bytes_per_launch": alg}}
    print(json.dumps(line), flush=True)
    del e

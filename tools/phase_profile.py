#!/usr/bin/env python
"""Phase breakdown of the dense sweep kernel (BRR_GPU_PROFILE=1): cycles of CTA 0 per phase, per block.
usage: python tools/phase_profile.py [n_markers] [sweeps] [block]"""
import os, sys
os.environ["BRR_GPU_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayesrrcmd_b200 import Engine

M = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
block = int(sys.argv[3]) if len(sys.argv) > 3 else 0
N = int(os.environ.get("BRR_BENCH_N", 500000))
pmiss = float(os.environ.get("BRR_PMISS", 0.001))
K = 4
e = Engine(N, M, K, 1, block_markers=block)
e.synth_dense(20260101, 0.01, 0.5, False, pmiss)
e.finalize(); e.set_model(np.array([[0.01, 0.001, 0.0001]]))
rng = np.random.default_rng(1)
y = rng.standard_normal(N); y = (y - y.mean()) / y.std(ddof=1)
e.set_state(eps=y, beta=np.zeros(M), comp=np.zeros(M), acum=np.zeros(M))
pi = np.full((1, K), 0.5 / K); pi[0, 0] = 0.5
sigmaG = np.array([0.5]); sigmaE = 0.5
info = e.info()
print("N=%d M=%d block=%d ctas=%d slice=%d" % (N, M, info["block_markers"], info["n_ctas"], info["slice_inds"]))
for s in range(sweeps):
    rand_uz = np.c_[rng.random(M), rng.standard_normal(M)]
    order = rng.permutation(M).astype(np.uint32)
    v, bs, out = e.sweep(order, rand_uz, pi, sigmaG, sigmaE)
    pc = e.profile_cycles(reset=True)
    nb = max(1, out.n_blocks)
    skip = set(e.PROFILE_COUNTERS) | set(e.PROFILE_OVERLAPPED) | {"prologue"}
    tot = sum(v_ for k, v_ in pc.items() if k not in skip)
    print("sweep %d: %.3f ms, %.2f us/block, nonskipped %.3f%%, blocks updated %.1f%%, walk iters/block %.2f, gram rounds/block %.3f, gram rows/block %.3f" % (
        s, out.kernel_ms, 1e3 * out.kernel_ms / nb, 100.0 * out.n_nonskipped / M, 100.0 * pc["blocks_updated"] / nb,
        pc["walk_iters"] / nb, pc["lazy_rounds"] / nb, pc["lazy_rows"] / nb))
    print("   walker cycles/block: " + "  ".join("%s=%d" % (k, pc[k] // nb) for k in e.PROFILE_PHASES if k not in skip)
          + "  | total/block=%d  prologue=%d" % (tot // nb, pc["prologue"]))
    print("   dot warps (overlapped) cycles/block: " + "  ".join("%s=%d" % (k, pc[k] // nb) for k in e.PROFILE_OVERLAPPED))
    m0 = v[0, 1:].sum()
    pi = rng.dirichlet(v[0] + 1.0)[None, :]
    sigmaG = np.array([max(1e-6, (bs[0] * 1.0 + 1e-8) / max(1.0, m0))])

"""Timing experiment: the dense sweep on a chain that never leaves beta = 0 (pure-noise phenotype, a prior that keeps everything out
of the model): no Gram rows, no restarts, no residual updates -- what remains is the streaming pipeline itself."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesrrcmd_b200 import Engine
N, M = 500000, int(os.environ.get("BRR_BENCH_M", 400000))
e = Engine(N, M, 4, 1, layout="dense")
e.synth_dense(20260101, 0.01, 0.5, False, 0.001)
e.finalize(); e.set_model(np.array([[0.01, 0.001, 0.0001]]))
rng = np.random.default_rng(1)
y = rng.standard_normal(N); y = (y - y.mean()) / y.std(ddof=1)
e.set_state(eps=y, beta=np.zeros(M), comp=np.zeros(M), acum=np.zeros(M))
pi = np.array([[1.0 - 3e-12, 1e-12, 1e-12, 1e-12]]); sigmaG = np.array([0.5])
ms = []
for it in range(6):
    order = rng.permutation(M).astype(np.uint32); uz = np.c_[rng.random(M), rng.standard_normal(M)]
    v, bs, out = e.sweep(order, uz, pi, sigmaG, 0.9)
    ms.append(out.kernel_ms)
    assert out.n_nonskipped == 0
k, r = e.last_sweep_info()
if os.environ.get("BRR_GPU_PROFILE") == "1" and k == 6:
    pc = e.profile_cycles6(reset=True)
    nblk = 6 * (M / 64)
    sys.stderr.write("v6 quiet chain, cycles per block: " + "  ".join("%s=%d" % (kk, v / nblk) for kk, v in pc.items() if not kk.startswith("w_") and kk != "c_iters" and v > nblk) + "  control-warp passes per block=%.2f\n" % (pc["c_iters"] / nblk))
    a = e.profile_ctas6(reset=True).astype(np.float64)
    per = a[:, :6] / np.maximum(1.0, a[:, 6:7])                        # cycles per block, per CTA
    names = e.PROFILE6_CTA[:6]
    sys.stderr.write("per CTA (first dot warp), cycles per block, min / median / max over %d CTAs: " % len(a)
                     + "  ".join("%s=%d/%d/%d" % (n, per[:, i].min(), np.median(per[:, i]), per[:, i].max()) for i, n in enumerate(names)) + "\n")
    o = np.argsort(per[:, 0])
    sys.stderr.write("the 12 CTAs that wait least for decisions (cta:smid dec_wait rows_wait mma_loop acc_wait): "
                     + "  ".join("%d:%d %d %d %d %d" % (c, a[c, 7], per[c, 0], per[c, 2], per[c, 3], per[c, 4]) for c in o[:12]) + "\n")
    sys.stderr.write("the 6 that wait most: " + "  ".join("%d:%d %d %d %d %d" % (c, a[c, 7], per[c, 0], per[c, 2], per[c, 3], per[c, 4]) for c in o[-6:]) + "\n")
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    np.save(os.path.join(out_dir, "quiet_ctas6.npy"), a)
print(json.dumps({"kernel": k, "dbg": os.environ.get("BRR_GPU_V5_DBG"), "ms_per_sweep": float(np.mean(ms[2:])), "us_per_block": 1e3 * float(np.mean(ms[2:])) / (M / 64), "M": M}))

"""Debug aid: N = 500k synthetic store, the same two sweeps through version 4 and version 6; first differences in sweep order."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesrrcmd_b200 import Engine
N, M, K = int(os.environ.get("DBG_N", 500000)), int(os.environ.get("DBG_M", 512)), 4
rng = np.random.default_rng(123)
S = np.array([[0.01, 0.001, 0.0001]])
y = rng.standard_normal(N); y = (y - y.mean()) / y.std(ddof=1)
p0 = float(os.environ.get("DBG_P0", 0.7))
pi = np.array([[p0, (1 - p0) / 3, (1 - p0) / 3, (1 - p0) / 3]]); sigmaG = np.array([0.3]); sigmaE = 0.7
sweeps = [(rng.permutation(M).astype(np.uint32), np.c_[rng.random(M), rng.standard_normal(M)]) for _ in range(3)]
def run(kern):
    os.environ["BRR_GPU_DENSE_KERNEL"] = kern
    e = Engine(N, M, K, 1)
    e.synth_dense(7, 0.01, 0.5, False, 0.001)
    e.finalize(); e.set_model(S)
    e.set_state(eps=y, beta=np.zeros(M), comp=np.zeros(M), acum=np.zeros(M))
    hist = []
    for order, uz in sweeps:
        v, bs, out = e.sweep(order, uz, pi, sigmaG, sigmaE)
        eps, b, comp, acum = e.get_state()
        hist.append((eps.copy(), b.copy(), comp.copy(), out, e.last_sweep_info()))
    return hist
h4 = run("v4"); h6 = run("v6")
for i, ((e4, b4, c4, o4, i4), (e6, b6, c6, o6, i6)) in enumerate(zip(h4, h6)):
    order = sweeps[i][0]
    db = np.abs(b4 - b6)[order]; bad = np.nonzero(db > 1e-12)[0]
    print("sweep", i, "v4", i4, o4.n_blocks, o4.n_nonskipped, "v6", i6, o6.n_blocks, o6.n_nonskipped, "max dbeta %.3e deps %.3e" % (db.max(), np.abs(e4 - e6).max()),
          "first bad positions", bad[:10].tolist(), "n bad", len(bad), "comp equal", np.array_equal(c4, c6))
    b_prev = h4[i - 1][1] if i else np.zeros(M)
    for p in bad[:4]:
        j = order[p]
        print("    pos", p, "marker", j, "beta_old", b_prev[j], "v4", b4[j], "v6", b6[j], " in-model in the 600 positions before:",
              np.nonzero(b_prev[order[max(0, p - 600):p]] != 0)[0].tolist()[-6:], "entered before:", np.nonzero((b_prev[order[max(0, p - 600):p]] == 0) & (b4[order[max(0, p - 600):p]] != 0))[0].tolist()[-6:], "of", min(p, 600))
    de = np.abs(e4 - e6)
    if de.max() > 1e-9:
        badi = np.nonzero(de > 1e-9)[0]
        print("    eps differs at", len(badi), "individuals; slices (3392):", np.unique(badi // 3392)[:20].tolist(), "...")

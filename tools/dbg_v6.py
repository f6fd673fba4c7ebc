"""Debug aid: one random problem of tests/test_gpu_dense.py through a forced dense kernel, errors against the C oracle listed in sweep order;
the same sweep repeated from the same state to see whether the result is deterministic."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_lib as O
from bayesrrcmd_b200 import Engine
from test_gpu_dense import _random_problem
N, max_ctas, M = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 200
p_miss, nG, K = 0.02, 2, 4
O.build()
pr = _random_problem(N * 7 + M, N, M, p_miss, nG, K)
rng = pr["rng"]
e = Engine(N, M, K, nG, block_markers=64, max_ctas=max_ctas)
e.upload_bed(pr["bed"]); e.finalize(); e.set_model(pr["S"], pr["groups"])
print(e.info())
beta = np.where(rng.random(M) < 0.3, rng.normal(size=M) * 0.05, 0.0)
comp = (beta != 0).astype(np.float64); acum = rng.random(M)
eps_o = pr["eps"].copy(); beta_o = beta.copy(); comp_o = comp.copy(); acum_o = acum.copy()
for sweep_i in range(2):
    order = rng.permutation(M).astype(np.uint32)
    if sweep_i == 1: order = order[: M - 7]
    rand_uz = np.c_[rng.random(M), rng.normal(size=M)]
    mdl = O.ModelArrays(N, pr["S"], pr["pi"], pr["sigmaG"], pr["sigmaE"])
    b_in = beta_o.copy(); e_in = eps_o.copy(); c_in = comp_o.copy(); a_in = acum_o.copy()
    v_o, bs_o, _ = O.sweep("dense", pr["bed"], M, order, rand_uz, mdl, pr["groups"], eps_o, beta_o, comp_o, acum_o, tokens=0)
    res = []
    for rep in range(3):
        e.set_state(eps=e_in, beta=b_in, comp=c_in, acum=a_in)
        v, bs, out = e.sweep(order, rand_uz, pr["pi"], pr["sigmaG"], pr["sigmaE"], 0.0, 0)
        eps, b, c, a = e.get_state()
        res.append((eps.copy(), b.copy()))
        err = np.abs(b - beta_o)[order]
        bad = np.nonzero(err > 1e-11)[0]
        print("sweep", sweep_i, "rep", rep, "kernel/restarts", e.last_sweep_info(), "blocks", out.n_blocks, "nonskipped", out.n_nonskipped,
              "max beta err %.3e eps err %.3e" % (err.max(), np.abs(eps - eps_o).max()), "bad positions", bad[:12].tolist())
        if rep == 0:
            for p in bad[:6]:
                j = order[p]
                print("   pos", p, "marker", j, "beta_in", b_in[j], "oracle", beta_o[j], "gpu", b[j], "in-model before it:", int((b_in[order[:p]] != 0).sum()),
                      "entered before it:", int(((b_in[order[:p]] == 0) & (beta_o[order[:p]] != 0)).sum()))
    print("   deterministic:", all(np.array_equal(res[0][0], r[0]) and np.array_equal(res[0][1], r[1]) for r in res[1:]))

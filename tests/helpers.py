"""Shared test helpers: golden fixtures (generated from the unmodified reference by
oracle/make_golden.py) and the replay loop both the oracle tests and the GPU tests use."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        self.N = int(self.z["N"]); self.M = int(self.z["M"])
        self.S = self.z["S"]
        self.K = self.S.shape[1] + 1
        self.nG = self.S.shape[0]
        self.y = self.z["y"]
        self.groups = self.z["groups"] if "groups" in self.z.files else None
        if "bed" in self.z.files:
            self.bed = self.z["bed"]
        else:
            self.bed = np.load(os.path.join(GOLDEN, str(self.z["bed_from"]) + ".npz"))["bed"]

    def chain(self, layout):
        pre = layout + "__"
        d = {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}
        if "inputs_shared_with_dense" in d:
            d["order"] = self.z["dense__order"]; d["rand_uz"] = self.z["dense__rand_uz"]
        return d

    def tier1(self):
        pre = "tier1__"
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}


def num_tol(num_ref, x_norm, eps_norm):
    """Tier-1 gate: |d| <= 1e-10*|num_ref| + 1e-13*||x||*||eps||."""
    return 1e-10 * abs(num_ref) + 1e-13 * x_norm * eps_norm


def replay(chain, iters, sweep_fn, get_state_fn, set_eps_fn, sparse=False):
    """Drive `iters` sweeps with the recorded inputs of a golden chain.

    sweep_fn(it, order, rand_uz, pi, sigmaG, sigmaE, eps_sum) -> (v, betasq, eps_sum_out)
    get_state_fn() -> (eps, beta, comp, acum);  set_eps_fn(eps) installs the residual.
    Between sweeps the host part of runGibbs is replayed from the recorded scalars: the residual is
    shifted by mu exactly as updateMu does ((eps + mu_old) - mu_new).  Yields per-iteration results.
    """
    mu = chain["mu"]
    for it in range(iters):
        if it == 0:
            set_eps_fn(chain["eps0"])
        else:
            eps, _, _, _ = get_state_fn()
            set_eps_fn((eps + mu[it - 1]) - mu[it])
        v, bs, es = sweep_fn(it, chain["order"][it].astype(np.uint32), chain["rand_uz"][it], chain["pi_in"][it],
                             chain["sigmaG_in"][it], float(chain["sigmaE_in"][it]), float(chain["epssum_in"][it]))
        eps, beta, comp, acum = get_state_fn()
        yield it, dict(v=v, betasq=bs, eps_sum=es, eps=eps, beta=beta, comp=comp, acum=acum)


def check_iteration(chain, it, res, tol=1e-8, check_acum=True):
    """Tier-2 gate: beta, components, acum, eps, v, betasqnG within `tol` of the reference chain."""
    def close(a, b, what):
        a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
        err = np.max(np.abs(a - b) / (1.0 + np.abs(b))) if a.size else 0.0
        assert err <= tol, "%s differs at iteration %d: %.3e" % (what, it, err)

    close(res["v"], chain["v_out"][it], "v")
    close(res["betasq"], chain["betasq_out"][it], "betasqnG")
    close(res["beta"].sum(), chain["chk_beta"][it], "sum(beta)")
    close((res["beta"] ** 2).sum(), chain["chk_beta2"][it], "sum(beta^2)")
    close(res["comp"].sum(), chain["chk_comp"][it], "sum(components)")
    if check_acum:
        close(res["acum"].sum(), chain["chk_acum"][it], "sum(acum)")
    close((res["eps"] ** 2).sum(), chain["chk_eps2"][it], "||eps||^2")
    full = list(chain["full_iters"])
    if it in full:
        k = full.index(it)
        close(res["beta"], chain["beta_full"][k], "beta")
        assert (res["comp"] == chain["comp_full"][k]).all(), "components differ at iteration %d" % it
        if check_acum:
            close(res["acum"], chain["acum_full"][k], "acum")
        close(res["eps"], chain["eps_full"][k], "eps")


def run_host_chain(N, M, S, y, iters, seed, sweep_fn, add_scalar_fn, eps_sum_fn=None):
    """The host part of BayesRBase::runGibbs (reference src/BayesRBase.cpp:186-253, one group) around a sweep callback, with numpy's
    generator for the draws: mu, the per-marker (u, z) table, the marker order, sigmaE, sigmaG, pi.  Used for STATISTICAL comparisons
    (posterior means), not for replay parity.

    sweep_fn(order, rand_uz, pi, sigmaG, sigmaE) -> (v[K], betasq, eps_sqnorm, n_nonzero);  add_scalar_fn(c) -> sum(eps) after eps += c
    Returns per-iteration arrays: mu, sigmaE, sigmaG, betasq, nonzero."""
    rng = np.random.default_rng(seed)
    K = S.shape[1] + 1
    pi = np.full((1, K), 0.5 / K); pi[0, 0] = 0.5
    sigmaG = np.array([rng.random()])
    sigmaE = float((y ** 2).sum()) / N * 0.5
    mu = 0.0
    v0 = s0 = 1e-4
    out = {k: np.zeros(iters) for k in ("mu", "sigmaE", "sigmaG", "betasq", "nonzero")}

    def inv_scaled_chisq(dof, scale):
        return 1.0 / rng.gamma(0.5 * dof, 1.0 / (0.5 * dof * scale))

    for it in range(iters):
        s = add_scalar_fn(mu)
        mu = rng.normal(s / N, np.sqrt(sigmaE / N))
        add_scalar_fn(-mu)
        rand_uz = np.c_[rng.random(M), rng.standard_normal(M)]
        order = rng.permutation(M).astype(np.uint32)
        v, bs, sq, nz = sweep_fn(order, rand_uz, pi, sigmaG, sigmaE)
        sigmaE = inv_scaled_chisq(v0 + N, (sq + v0 * s0) / (v0 + N))
        m0 = v[1:].sum()
        sigmaG = np.array([inv_scaled_chisq(v0 + m0, (bs * m0 + v0 * s0) / (v0 + m0))])
        pi = rng.dirichlet(v + 1.0)[None, :]
        out["mu"][it] = mu; out["sigmaE"][it] = sigmaE; out["sigmaG"][it] = sigmaG[0]; out["betasq"][it] = bs; out["nonzero"][it] = nz
    return out

"""GPU: the drop-in boundary, compiled.  oracle/gpu_graph_ref.cpp is INTEGRATION.md's GpuSweepGraph -- an AnalysisGraph
(reference src/analysisgraph.hpp:9-26) whose exec() calls brr_gpu_sweep -- compiled against the reference's own headers into
oracle/_ref/libbrr_ref_gpu.so.  Here the UNMODIFIED reference driver (BayesRBase::runGibbs, src/BayesRBase.cpp:130-285: its init,
its RNG, updateMu, shuffle, sigmaE / sigmaG / pi draws) runs the same seeded chain twice: once through the reference's own
sequential marker loop, once through the GPU graph.  The only non-reference code between runGibbs and the kernel is that graph class.
The chains must agree within 1e-8 (tier 2 of BASELINE.json) -- every iteration, from the seed."""
import os

import numpy as np
import pytest

import ref_lib as R
from helpers import Golden

pytestmark = pytest.mark.gpu


def _close(a, b, what, tol=1e-8):
    err = np.max(np.abs(a - b) / (1.0 + np.abs(b)))
    assert err <= tol, "%s: %.3e" % (what, err)


@pytest.mark.parametrize("name,layout,iters", [("uk10k", 0, 10), ("groups", 0, 10), ("missing_small", 0, 10), ("uk10k", 1, 6), ("uk10k", 2, 6)])
def test_reference_rungibbs_through_the_gpu_graph(name, layout, iters):
    if not (R.ref_available() and os.path.exists(R._path("gpu"))):
        pytest.skip("oracle/_ref is not built")
    g = Golden(name)
    y = g.y.astype(np.float32)
    ref, _ = R.run_chain(g.bed, g.N, g.M, y, g.S, seed=7, iterations=iters, layout=layout, groups=g.groups, graph="sequential")
    gpu, _ = R.run_chain(g.bed, g.N, g.M, y, g.S, seed=7, iterations=iters, layout=layout, groups=g.groups, graph="gpu")
    assert np.array_equal(ref["order"], gpu["order"]) and np.array_equal(ref["rand_uz"][0], gpu["rand_uz"][0])
    for it in range(iters):
        assert np.array_equal(ref["comp_out"][it], gpu["comp_out"][it]), "components differ at iteration %d" % it
        assert np.array_equal(ref["v_out"][it], gpu["v_out"][it])
        for k in ("beta_out", "eps_out", "acum_out", "betasq_out", "sigmaE_post", "sigmaG_post", "pi_post", "mu"):
            _close(gpu[k][it], ref[k][it], "%s at iteration %d" % (k, it))
    assert (ref["beta_out"][-1] != 0).sum() > 0

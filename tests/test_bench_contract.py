"""CPU: the JSON line of bench.py's reference arm (the reference's own CPU code on a bounded sample, small sizes here) carries the keys the
driver pairs with the GPU arm's line, and the GPU arm refuses to run without a GPU (no CPU fallback behind the headline number)."""
import json
import os
import subprocess
import sys

import pytest

import ref_lib as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, cwd=ROOT)


@pytest.mark.skipif(not R.ref_available(fast=True), reason="oracle/_ref is not built")
def test_reference_arm_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--n-inds", "20000", "--n-markers", "2000", "--cpu-markers", "64"])
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    for k in ("metric", "value", "unit", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["unit"] == "marker updates/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["e2e"]["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same metric string as the GPU arm prints (the driver pairs the two lines by it)
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"metric": METRIC') >= 2 and d["metric"] in src


def test_gpu_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1", "--warmup", "0", "--n-inds", "20000", "--n-markers", "2000", "--no-cpu-baseline"])
    assert r.returncode != 0 and not r.stdout.strip().startswith("{")

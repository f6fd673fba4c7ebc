"""CPU: the walk's branch-free exponential (bayesrrcmd_b200/csrc/sampler.cuh: exp_nb, shared host/device code) against libm over
the whole argument range the sampler can feed it (the reference's 700 guard)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exp_nb_against_libm(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "check_exp_nb")
    subprocess.check_call([nvcc, "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tools", "check_exp_nb.cu")])
    out = subprocess.check_output([exe]).decode()
    worst = float(re.search(r"worst rel err ([0-9.e+-]+)", out).group(1))
    assert worst < 1e-15, out

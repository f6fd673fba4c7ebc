"""The count-carrying accumulator words of the cross-GPU per-block round (sweep_dense.cuh: kCnt1 / line_count / line_value):
host-side property test of the encoding, compiled from the product header with nvcc (runs on the CPU)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_count_words_decode(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "count_words")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-arch=sm_100a", "-I", os.path.join(ROOT, "bayesrrcmd_b200", "csrc"), "-o", exe,
                           os.path.join(ROOT, "tests", "native", "count_words.cu")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "bad 0" in out.stdout

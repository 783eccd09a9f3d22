"""Device-side unit checks that need their own tiny CUDA program (built here with nvcc)."""
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shared_divisor_is_bit_identical_to_ieee_division(cuda_device, tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "div_check")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-fmad=false", "-std=c++17",
                           "-I", os.path.join(ROOT, "centroidalplanner_b200", "csrc"),
                           os.path.join(ROOT, "tests", "native", "div_check.cu"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatches 0" in r.stdout

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


def test_ifopt_surface_views_match_oracle_bit_for_bit(cuda_device, tmp_path):
    """cplb/ifopt_views.hpp (VariableSet / ConstraintSet / CostTerm views over the batched buffers) driven through
    ifopt::Problem's Evaluate* calls; ifopt/Eigen are the stand-in headers of oracle/refshim."""
    gxx = shutil.which("g++") or "g++"
    exe = str(tmp_path / "ifopt_views_check")
    pkg = os.path.join(ROOT, "centroidalplanner_b200")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libcpl_oracle.so"], stdout=subprocess.DEVNULL)
    subprocess.check_call([gxx, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(pkg, "cpp", "include"),
                           "-I", os.path.join(ROOT, "oracle", "refshim"), "-I", os.path.join(ROOT, "oracle"),
                           os.path.join(ROOT, "tests", "native", "ifopt_views_check.cpp"), "-o", exe,
                           "-L", pkg, "-lcplb", "-L", os.path.join(ROOT, "oracle"), "-lcpl_oracle",
                           f"-Wl,-rpath,{pkg}", f"-Wl,-rpath,{os.path.join(ROOT, 'oracle')}"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failures" in r.stdout

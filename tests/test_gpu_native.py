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


def _build_native(tmp_path, src, extra=()):
    gxx = shutil.which("g++") or "g++"
    exe = str(tmp_path / os.path.splitext(src)[0])
    pkg = os.path.join(ROOT, "centroidalplanner_b200")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libcpl_oracle.so"], stdout=subprocess.DEVNULL)
    subprocess.check_call([gxx, "-std=c++17", "-O1", "-pthread", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(pkg, "cpp", "include"),
                           "-I", os.path.join(ROOT, "oracle", "refshim"), "-I", os.path.join(ROOT, "oracle"),
                           os.path.join(ROOT, "tests", "native", src), "-o", exe, "-L", pkg, "-lcplb",
                           "-L", os.path.join(ROOT, "oracle"), "-lcpl_oracle", f"-Wl,-rpath,{pkg}",
                           f"-Wl,-rpath,{os.path.join(ROOT, 'oracle')}", *extra])
    return exe


def test_lock_step_solver_threads_cost_one_launch_per_round(cuda_device, tmp_path):
    """96 solver threads, each driving its own ifopt::Problem view through IPOPT's four callbacks for 12 rounds:
    exactly 12 batched evaluations, every value bit-identical to the CPU replay (tests/native/lockstep_check.cpp)."""
    exe = _build_native(tmp_path, "lockstep_check.cpp")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "-> 12 batched evaluations, 0 mismatches" in r.stdout


def test_pybind11_module_matches_ctypes_path(cuda_device):
    """bindings/python/pycplb.cpp forwards to the same C ABI: same bits as the ctypes mirror."""
    import sys

    import numpy as np

    subprocess.check_call(["make", "-C", os.path.join(ROOT, "bindings", "python")], stdout=subprocess.DEVNULL)
    sys.path.insert(0, os.path.join(ROOT, "centroidalplanner_b200"))
    import pycplb

    import centroidalplanner_b200 as cpl
    from centroidalplanner_b200 import synthetic

    x = synthetic.superquadric_batch(3000)
    sq = synthetic.SUPERQUADRIC
    e1 = pycplb.Superquadric()
    e1.SetParameters(sq["C"], sq["R"], sq["P"])
    e1.SetMu(0.5)
    p1 = pycplb.BatchedProblem(synthetic.NAMES4, 100.0, e1)
    p1.SetManipulationWrench(synthetic.TESTBASIC["wrench"])
    p1.SetForceThreshold("contact2", 7.0)
    e2 = cpl.Superquadric()
    e2.SetParameters(sq["C"], sq["R"], sq["P"])
    e2.SetMu(0.5)
    p2 = cpl.BatchedCplProblem(synthetic.NAMES4, 100.0, e2)
    p2.SetManipulationWrench(synthetic.TESTBASIC["wrench"])
    p2.SetForceThreshold("contact2", 7.0)
    a = p1.eval(x, g=True, jac=True, cost=True, grad=True)
    b = p2.eval(x, g=True, jac=True, cost=True, grad=True)
    for k in ("g", "jac", "cost", "grad"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    r1, c1 = p1.GetJacobianStructure()
    r2, c2 = p2.GetJacobianStructure()
    assert np.array_equal(r1, r2) and np.array_equal(c1, c2)


def test_pybind11_module_round_two_entry_points(cuda_device):
    """The pybind11 module forwards the round-2 ABI too: sharded problems, packed Jacobian slices, native lock-step solves."""
    import sys

    import numpy as np
    import torch

    subprocess.check_call(["make", "-C", os.path.join(ROOT, "bindings", "python")], stdout=subprocess.DEVNULL)
    sys.path.insert(0, os.path.join(ROOT, "centroidalplanner_b200"))
    import pycplb

    from centroidalplanner_b200 import synthetic

    def ground():
        e = pycplb.Ground()
        e.SetGroundZ(0.1)
        e.SetMu(0.5)
        return e

    x = synthetic.ground_batch(5000)
    single = pycplb.BatchedProblem(synthetic.NAMES4, 100.0, ground())
    sharded = pycplb.BatchedProblem(synthetic.NAMES4, 100.0, ground(), devices=[0, 0, 0])
    assert sharded.GetNumberOfShards() == 3 and sharded.GetShard(1, 5000)[1] % 32 == 0
    a = single.eval(x, g=True, jac=True)
    b = sharded.eval(x, g=True, jac=True, jac_packed=True)
    pmap = single.GetPackedJacobianMap()
    assert np.array_equal(a["g"], b["g"]) and np.array_equal(a["jac"][:, pmap], b["jac"])
    kind, source = single.GetJacobianSlotSources()
    c = sharded.eval(x, g=True, jac=True, jac_packed=2)          # CPLB_JAC_COMPUTED slices
    assert np.array_equal(a["jac"][:, kind == 3], c["jac"]) and np.array_equal(a["g"], c["g"])
    neg = np.nonzero(kind == 2)[0]
    assert len(neg) + int((kind == 1).sum()) == 48 and np.array_equal(a["jac"][:, neg], -x[:, source[neg]])
    # native solve through the module: TestBasic's ground problem, equilibrium lines on every instance
    single.SetCoMWeight(2.0)
    single.SetForceWeight(0.0)
    single.SetManipulationWrench([100.0, 0, 0, 0, 0, 100.0])
    for nm in synthetic.NAMES4:
        single.SetPosBounds(nm, [-0.3, -0.3, 0.0], [0.3, 0.3, 1.0])
    N, n = 32, single.n
    rng = np.random.default_rng(1)
    x0 = np.zeros((N, n))
    x0[:, 2] = 0.5
    for k in range(4):
        x0[:, 3 + 9 * k:6 + 9 * k] = [1.0 + 0.5 * k, -1.0 + 0.25 * k, 245.25]
        x0[:, 6 + 9 * k:9 + 9 * k] = [0.24 * (1 if k in (0, 3) else -1), 0.24 * (1 if k < 2 else -1), 0.5]
        x0[:, 9 + 9 * k:12 + 9 * k] = [0.0, 0.0, 1.0]
    x0 += rng.normal(0, 0.02, x0.shape)
    d = lambda *shape, dt=torch.float64: torch.empty(*shape, dtype=dt, device=cuda_device)  # noqa: E731
    xd, xs, st, it, cost, viol, dual = torch.from_numpy(x0).to(cuda_device), d(N, n), d(N, dt=torch.int32), d(N, dt=torch.int32), d(N), d(N), d(N)
    rounds, evals, _ = single.solve_device(N, xd.data_ptr(), xs.data_ptr(), st.data_ptr(), it.data_ptr(), cost.data_ptr(), viol.data_ptr(), dual.data_ptr())
    assert rounds >= 0 and evals == 1 + 4 * rounds and st.tolist() == [0] * N     # 32 instances: the tail may take them from round 0
    sol = xs.cpu().numpy()
    F = sol[:, 3:].reshape(N, 4, 9)[:, :, 0:3].sum(axis=1)
    assert np.abs(F - [100.0, 0.0, 981.0]).max() < 1e-6

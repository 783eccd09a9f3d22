"""Randomised parity sweep: random problem shapes (1..10 contacts, random names in random order, every environment kind),
random parameters of every kind the ABI has, inputs over many orders of magnitude and -- for the pow-free problems --
special values (zeros, subnormals, 1e150, +-inf, NaN) sprinkled over the batch.  Same bar as tests/test_gpu_parity.py:
bit-exact where no pow() is upstream, <= 1e-12 relative elsewhere, NaN/Inf matched by position."""
import string

import numpy as np
import pytest

import centroidalplanner_b200 as cpl
from centroidalplanner_b200 import synthetic

from helpers import OracleProblem, assert_parity
from test_gpu_parity import LAYOUTS, run_device

pytestmark = pytest.mark.gpu


def random_problem(rng):
    nc = int(rng.integers(1, 11))
    names = set()
    while len(names) < nc:
        names.add("".join(rng.choice(list(string.ascii_lowercase + "_0123456789"), size=int(rng.integers(1, 12)))))
    names = list(names)
    rng.shuffle(names)
    env_name = ["none", "ground", "superquadric"][int(rng.integers(0, 3))]
    mass = float(rng.uniform(1.0, 500.0))
    env = {"none": None, "ground": cpl.Ground, "superquadric": cpl.Superquadric}[env_name]
    env = env() if env is not None else None
    prob = cpl.BatchedCplProblem(names, mass, env)
    op = OracleProblem(names, env_name, mass)
    mu = float(rng.uniform(0.05, 1.5))
    wrench = rng.normal(0.0, 100.0, 6)
    sq = None
    if env_name == "superquadric":
        integer_P = bool(rng.integers(0, 2))
        P = rng.integers(2, 13, 3).astype(float) if integer_P else rng.uniform(2.0, 6.0, 3)
        sq = dict(C=rng.uniform(-1.0, 1.0, 3), R=rng.uniform(0.2, 2.0, 3), P=P, integer_P=integer_P)
    for target, e in ((prob, env), (op, op)):
        r = np.random.default_rng(12345)        # identical parameter draws on both sides
        if env_name == "ground":
            e.SetGroundZ(float(r.uniform(-1, 1)))
        if sq is not None:
            e.SetParameters(sq["C"], sq["R"], sq["P"])
        (e if env_name != "none" else target).SetMu(mu)
        target.SetManipulationWrench(wrench)
        target.SetCoMRef(r.uniform(-1, 1, 3))
        target.SetCoMWeight(float(r.uniform(0, 5)))
        for nm in names:
            target.SetForceThreshold(nm, float(r.uniform(0, 50)))
            target.SetPosRef(nm, r.uniform(-1, 1, 3))
            target.SetForceRef(nm, r.uniform(-300, 300, 3))
            target.SetContactPosWeight(nm, float(r.uniform(0, 3)))
            target.SetContactForceWeight(nm, float(r.uniform(0, 1e-2)))
    return prob, op.o, names, env_name, sq


@pytest.mark.parametrize("seed", range(24))
def test_random_problem_matches_oracle(seed, cuda_device):
    rng = np.random.default_rng(1000 + seed)
    prob, o, names, env_name, sq = random_problem(rng)
    nc, N = len(names), 515
    if env_name == "superquadric":
        x = synthetic.superquadric_batch(N, nc, 7000 + seed)
        for k in range(nc):                      # positions C +- U(0.05, 2) R; positive side only for fractional curvatures
            sgn = rng.choice([-1.0, 1.0], size=(N, 3)) if sq["integer_P"] else 1.0
            x[:, 3 + 9 * k + 3:3 + 9 * k + 6] = sq["C"] + sgn * rng.uniform(0.05, 2.0, (N, 3)) * sq["R"]
    else:
        x = synthetic.ground_batch(N, nc, 7000 + seed)
    x *= 10.0 ** rng.uniform(-3, 3, (N, 1)) if env_name != "superquadric" else 1.0     # whole instances rescaled
    if env_name != "superquadric":
        special = np.array([0.0, -0.0, 5e-324, 1e-310, 1e150, -1e150, np.inf, -np.inf, np.nan])
        hit = rng.random(x.shape) < 0.01
        x[hit] = rng.choice(special, size=int(hit.sum()))
    with np.errstate(all="ignore"):
        want = o.eval_batch(x, nthreads=4)
    for layout in LAYOUTS:
        got = run_device(prob, x, layout, cuda_device, g=True, jac=True, cost=True, grad=True)
        assert_parity(got, want, o, f"fuzz{seed}/{env_name}/nc{nc}/layout{layout}", x)
    r, c = prob.GetJacobianStructure()
    ro, co = o.structure()
    assert np.array_equal(r, ro) and np.array_equal(c, co)

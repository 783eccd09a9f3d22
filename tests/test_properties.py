"""Size-independent properties of the evaluation, checked on the oracle (CPU) and on the GPU path.

* relabelling: permuting the caller's contact-name vector permutes columns, not values -- rows follow the SORTED
  names (std::map order, src/CplProblem.cpp:42) and must not move;
* force-balance linearity: rows 0..2 are linear in the forces; doubling every F and the wrench and the mass doubles them;
* the friction rows are invariant under rotation of F and n about the z axis; the Ground rows do not depend on F."""
import numpy as np
import pytest

import centroidalplanner_b200 as cpl
from centroidalplanner_b200 import synthetic
from oracle import cpl_oracle_py as orc


def _oracle(names, env=orc.ENV_GROUND, mass=100.0, wrench=(100, 0, 0, 0, 0, 100)):
    o = orc.Oracle(names, env, mass)
    o.set_ground_z(0.1)
    o.set_mu(0.5)
    o.set_wrench(wrench)
    for nm in names:
        o.set_force_threshold(nm, 3.0 + len(nm))
    return o


def _permute_x(x, perm):
    """x for names in order `perm` (new position a holds old contact perm[a])."""
    nc = len(perm)
    out = x.copy()
    for a, k in enumerate(perm):
        out[:, 3 + 9 * a:12 + 9 * a] = x[:, 3 + 9 * k:12 + 9 * k]
    return out


@pytest.mark.parametrize("env", [orc.ENV_NONE, orc.ENV_GROUND, orc.ENV_SUPERQUADRIC])
def test_relabelling_contacts_moves_columns_not_values(env):
    names = ["d_foot", "a_foot", "c_hand", "b_hand"]
    rng = np.random.default_rng(5)
    gen = synthetic.superquadric_batch if env == orc.ENV_SUPERQUADRIC else synthetic.ground_batch
    x = gen(50, 4, 99)
    o1 = _oracle(names, env)
    sq = synthetic.SUPERQUADRIC
    o1.set_superquadric(sq["C"], sq["R"], sq["P"])
    e1 = o1.eval_batch(x)
    r1, c1 = o1.structure()
    J1 = np.zeros((50, o1.m, o1.n))
    J1[:, r1, c1] = e1["jac"]
    for _ in range(4):
        perm = rng.permutation(4)
        o2 = _oracle([names[k] for k in perm], env)
        o2.set_superquadric(sq["C"], sq["R"], sq["P"])
        e2 = o2.eval_batch(_permute_x(x, perm))
        r2, c2 = o2.structure()
        J2 = np.zeros((50, o2.m, o2.n))
        J2[:, r2, c2] = e2["jac"]
        assert np.array_equal(e1["g"], e2["g"], equal_nan=True)          # rows follow sorted names: unchanged, bit for bit
        assert np.array_equal(e1["cost"], e2["cost"])
        colmap = np.arange(o1.n)
        for a, k in enumerate(perm):
            colmap[3 + 9 * a:12 + 9 * a] = np.arange(3 + 9 * k, 12 + 9 * k)
        assert np.array_equal(J2, J1[:, :, colmap], equal_nan=True)      # columns follow the caller's vector
        assert np.array_equal(e2["grad"], e1["grad"][:, colmap])


def test_force_balance_rows_are_linear():
    names = synthetic.NAMES4
    x = synthetic.ground_batch(100)
    o1 = _oracle(names, mass=50.0, wrench=(10, -20, 30, 0, 0, 0))
    o2 = _oracle(names, mass=100.0, wrench=(20, -40, 60, 0, 0, 0))
    x2 = x.copy()
    for k in range(4):
        x2[:, 3 + 9 * k:6 + 9 * k] *= 2.0
    g1, g2 = o1.eval_batch(x, want=("g",))["g"], o2.eval_batch(x2, want=("g",))["g"]
    assert np.array_equal(2.0 * g1[:, :3], g2[:, :3])      # scaling by 2 is exact in binary floating point
    assert np.array_equal(2.0 * g1[:, 3:6], g2[:, 3:6])    # moments are linear in F too (wrench rows 3..5 are 0 here)


def test_friction_rows_are_invariant_under_rotation_about_z():
    names = synthetic.NAMES4
    x = synthetic.ground_batch(200)
    o = _oracle(names)
    th = 0.7
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    xr = x.copy()
    for k in range(4):
        xr[:, 3 + 9 * k:6 + 9 * k] = x[:, 3 + 9 * k:6 + 9 * k] @ R.T
        xr[:, 9 + 9 * k:12 + 9 * k] = x[:, 9 + 9 * k:12 + 9 * k] @ R.T
    g, gr = o.eval_batch(x, want=("g",))["g"], o.eval_batch(xr, want=("g",))["g"]
    for j in range(4):
        rows = [6 + 6 * j + 4, 6 + 6 * j + 5]
        assert np.abs(g[:, rows] - gr[:, rows]).max() <= 1e-10
        assert np.array_equal(g[:, 6 + 6 * j], gr[:, 6 + 6 * j])   # Ground value depends on p only


@pytest.mark.gpu
@pytest.mark.parametrize("layout", [cpl.INSTANCE_MAJOR, cpl.COMPONENT_MAJOR])
def test_gpu_relabelling_property(layout, cuda_device):
    """The same relabelling property on the CUDA path at 100,000 instances (no oracle involved): values must be
    bit-identical under any order of the caller's name vector."""
    import torch

    names = ["d_foot", "a_foot", "c_hand", "b_hand", "f_knee"]
    N = 100_000
    x = synthetic.ground_batch(N, 5, 123)

    def run(nm, xx):
        env = cpl.Ground()
        env.SetGroundZ(0.1)
        env.SetMu(0.5)
        prob = cpl.BatchedCplProblem(nm, 80.0, env)
        prob.SetManipulationWrench([1, 2, 3, 4, 5, 6])
        for s in nm:
            prob.SetForceThreshold(s, 3.0 + len(s))
        xd = torch.from_numpy(xx).to(cuda_device)
        if layout == cpl.COMPONENT_MAJOR:
            xd = xd.t().contiguous()
        out = prob.eval(xd, g=True, jac=True, cost=True, grad=True, layout=layout)
        torch.cuda.synchronize()
        r, c = prob.GetJacobianStructure()
        jac = out["jac"] if layout == cpl.INSTANCE_MAJOR else out["jac"].t()
        g = out["g"] if layout == cpl.INSTANCE_MAJOR else out["g"].t()
        return g.contiguous(), jac.contiguous(), out["cost"], r, c

    g1, j1, c1, r1, col1 = run(names, x)
    perm = np.array([3, 0, 4, 2, 1])
    g2, j2, c2, r2, col2 = run([names[k] for k in perm], _permute_x(x, perm))
    assert torch.equal(g1.view(torch.int64), g2.view(torch.int64)) and torch.equal(c1.view(torch.int64), c2.view(torch.int64))
    # slot-by-slot: entry (row, col) of problem 2 corresponds to (row, colmap[col]) of problem 1
    colmap = np.arange(3 + 9 * 5)
    for a, k in enumerate(perm):
        colmap[3 + 9 * a:12 + 9 * a] = np.arange(3 + 9 * k, 12 + 9 * k)
    key1 = {(int(r), int(c)): s for s, (r, c) in enumerate(zip(r1, col1))}
    order = torch.tensor([key1[(int(r), int(colmap[c]))] for r, c in zip(r2, col2)], device=cuda_device)
    assert torch.equal(j2.view(torch.int64), j1[:, order].contiguous().view(torch.int64))

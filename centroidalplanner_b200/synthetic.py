"""Synthetic contact/CoM batches for the BASELINE.json configs.

Pure NumPy, seeded; used by tests/ and bench.py so both sides of every comparison see the same
inputs.  Parameters are the reference's own test parameters (tests/TestBasic.cpp:64-99, :138-175).
"""
from __future__ import annotations

import numpy as np

NAMES4 = ["contact1", "contact2", "contact3", "contact4"]                       # TestBasic.cpp:70-74
# config 4: given "r_" first so that vector order != sorted order
NAMES8 = ["r_foot_a", "r_foot_b", "r_hand_a", "r_hand_b", "l_foot_a", "l_foot_b", "l_hand_a", "l_hand_b"]

TESTBASIC = dict(mass=100.0, ground_z=0.1, mu=0.5, W_CoM=2.0, W_F=0.0, wrench=[100.0, 0, 0, 0, 0, 100.0],
                 p_lb=[-0.3, -0.3, 0.0], p_ub=[0.3, 0.3, 1.0])                  # TestBasic.cpp:67-99
SUPERQUADRIC = dict(C=[0.0, 0.0, 1.0], R=[0.3, 0.3, 10.0], P=[10.0, 10.0, 10.0], mu=0.5,
                    p_lb=[-0.5, -0.5, 0.5], p_ub=[0.5, 0.5, 1.5])               # TestBasic.cpp:150-169

FEET = np.array([[0.3, 0.2, 0.1], [0.3, -0.2, 0.1], [-0.3, 0.2, 0.1], [-0.3, -0.2, 0.1]])
HANDS = np.array([[0.4, 0.3, 1.0], [0.4, -0.3, 1.0], [-0.4, 0.3, 1.0], [-0.4, -0.3, 1.0]])


def _pack(com, F, p, n):
    """x in column-map order: CoM, then per contact (vector order) F, p, n.  Shapes (N,3), (N,nc,3)."""
    N, nc = F.shape[0], F.shape[1]
    x = np.empty((N, 3 + 9 * nc))
    x[:, 0:3] = com
    blk = np.concatenate([F, p, n], axis=2)  # (N, nc, 9)
    x[:, 3:] = blk.reshape(N, 9 * nc)
    return x


def _common(rng, N, nc):
    com = rng.uniform([-0.2, -0.2, 0.8], [0.2, 0.2, 1.2], size=(N, 3))
    F = np.stack([rng.uniform(-50, 50, (N, nc)), rng.uniform(-50, 50, (N, nc)), rng.uniform(100, 400, (N, nc))], axis=2)
    n = np.array([0.0, 0.0, 1.0]) + 0.1 * rng.standard_normal((N, nc, 3))  # deliberately not unit
    return com, F, n


def ground_batch(N, nc=4, seed=1002):
    """configs 2 (nc=4, seed 1002) and 4 (nc=8, seed 1004): flat-ground instances, (N, n) instance-major."""
    rng = np.random.default_rng(seed)
    com, F, n = _common(rng, N, nc)
    anchors = FEET if nc <= 4 else np.concatenate([FEET, HANDS])
    anchors = np.resize(anchors, (nc, 3))
    p = anchors[None, :, :] + rng.uniform(-0.1, 0.1, (N, nc, 3))
    return _pack(com, F, p, n)


def superquadric_batch(N, nc=4, seed=1003):
    """config 3: contacts scattered around the superquadric, away from its 1/(p-C)^2 poles."""
    rng = np.random.default_rng(seed)
    com, F, n = _common(rng, N, nc)
    C = np.array(SUPERQUADRIC["C"])
    Rp = np.array([0.3, 0.3, 0.4])
    sgn = rng.choice([-1.0, 1.0], size=(N, nc, 3))
    p = C + sgn * rng.uniform(0.2, 1.2, (N, nc, 3)) * Rp
    return _pack(com, F, p, n)


def configure_testbasic(problem, names):
    """testGroundEnv parameters (tests/TestBasic.cpp:83-99) on a problem-like object (product or oracle adapter)."""
    problem.SetCoMWeight(TESTBASIC["W_CoM"])
    problem.SetForceWeight(TESTBASIC["W_F"])
    for nm in names:
        problem.SetPosBounds(nm, TESTBASIC["p_lb"], TESTBASIC["p_ub"])
    problem.SetManipulationWrench(TESTBASIC["wrench"])

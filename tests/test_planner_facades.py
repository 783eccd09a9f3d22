"""The planner facades' host logic (cpl::CentroidalPlanner / cpl::CoMPlanner), mirrored for the batched problem.
Setups follow tests/TestBasic.cpp; what the reference asserts after an IPOPT solve is asserted here on the
evaluator at a hand-built feasible point instead (IPOPT is absent)."""
import numpy as np
import pytest

import centroidalplanner_b200 as cpl

NAMES = ["contact1", "contact2", "contact3", "contact4"]


def test_centroidal_planner_validation():
    with pytest.raises(ValueError, match="Invalid robot mass"):
        cpl.BatchedCentroidalPlanner(NAMES, -1.0, cpl.Ground())
    pl = cpl.BatchedCentroidalPlanner(NAMES, 100.0, cpl.Ground())
    for call in (lambda: pl.SetForceBounds("nose", [0] * 3, [1] * 3), lambda: pl.GetPosRef("nose"),
                 lambda: pl.SetContactPosWeight("nose", 1.0), lambda: pl.SetForceThreshold("nose", 1.0)):
        with pytest.raises(ValueError, match="Invalid contact name: 'nose'"):   # std::invalid_argument at the facade level
            call()
    with pytest.raises(ValueError, match="Invalid weight"):
        pl.SetForceWeight(-1.0)
    with pytest.raises(ValueError, match="Invalid force threshold"):
        pl.SetForceThreshold("contact1", -5.0)
    pl.SetPosWeight(3.0)
    assert pl.GetPosWeight() == {nm: 3.0 for nm in NAMES}
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU evaluation path"):   # Solve() never falls back to a CPU evaluation
            pl.Solve()


def test_force_threshold_is_not_forwarded_for_a_zero_force_contact():
    """src/CentroidalPlanner.cpp:340."""
    pl = cpl.BatchedCentroidalPlanner(NAMES, 100.0, cpl.Ground())
    pl.SetForceThreshold("contact2", 20.0)
    assert pl.GetForceThreshold("contact2") == 20.0
    pl.SetForceBounds("contact2", np.zeros(3), np.zeros(3))
    pl.SetForceThreshold("contact2", 35.0)
    assert pl.GetForceThreshold("contact2") == 20.0          # ignored: the contact's force is pinned to zero


def test_com_planner_bookkeeping_follows_testcomplanner():
    """tests/TestBasic.cpp:225-292: contacts at (+-1, +-1, 0), mu 0.5, contact4 lifting, F_thr 20 everywhere."""
    pl = cpl.BatchedCoMPlanner(NAMES, 100.0)
    prob = pl.GetCplProblem()
    assert (prob.n, prob.m, prob.nnz) == (39, 14, 114)       # env == nullptr: FrictionCone rows only
    assert pl.GetPosWeight() == {nm: 0.0 for nm in NAMES} and pl.GetForceWeight() == {nm: 0.0 for nm in NAMES}
    pts = {"contact1": [1, 1, 0], "contact2": [1, -1, 0], "contact3": [-1, 1, 0], "contact4": [-1, -1, 0]}
    for nm, p in pts.items():
        assert list(pl.GetContactNormal(nm)) == [0, 0, 1]
        with pytest.raises(RuntimeError, match="not set"):
            pl.GetContactPosition(nm)
        pl.SetContactPosition(nm, p)
        assert list(pl.GetContactPosition(nm)) == p
    pl.SetMu(0.5)
    with pytest.raises(ValueError, match="Invalid friction coefficient"):
        pl.SetMu(0.0)
    for nm in NAMES:
        pl.SetForceThreshold(nm, 20.0)
    pl.SetLiftingContact("contact4")
    assert pl.GetLiftingContacts() == ["contact4"] and pl.GetForceThreshold("contact4") == 0.0
    lb, ub = prob.GetBoundsOnOptimizationVariables()
    c4 = prob.GetBlockColumn(cpl.BLOCK_FORCE, "contact4")
    assert (lb[c4:c4 + 3] == 0).all() and (ub[c4:c4 + 3] == 0).all()
    cp = prob.GetBlockColumn(cpl.BLOCK_POSITION, "contact2")
    assert list(lb[cp:cp + 3]) == [1, -1, 0] == list(ub[cp:cp + 3])
    with pytest.raises(RuntimeError, match="is not a lifting contact"):
        pl.ResetLiftingContact("contact1")
    pl.ResetLiftingContact("contact4")
    assert pl.GetLiftingContacts() == [] and pl.GetForceThreshold("contact4") == 20.0
    lb, ub = prob.GetBoundsOnOptimizationVariables()
    assert (lb[c4:c4 + 3] == -1e3).all() and (ub[c4:c4 + 3] == 1e3).all()


@pytest.mark.gpu
def test_com_planner_feasible_point_satisfies_what_testcomplanner_asserts(cuda_device):
    """At a hand-built static-balance point with contact4 lifting (tests/TestBasic.cpp:244-290): friction rows <= 0,
    Sigma F = (0, 0, m g), Sigma tau = 0 -- evaluated on the GPU for a batch of identical instances."""
    import torch

    pl = cpl.BatchedCoMPlanner(NAMES, 100.0)
    pl.SetMu(0.5)
    for nm in NAMES:
        pl.SetForceThreshold(nm, 20.0)
    pl.SetLiftingContact("contact4")
    prob = pl.GetCplProblem()
    pts = {"contact1": [1, 1, 0], "contact2": [1, -1, 0], "contact3": [-1, 1, 0], "contact4": [-1, -1, 0]}
    Fz = {"contact1": 40.0, "contact2": 470.5, "contact3": 470.5, "contact4": 0.0}   # sums to m g = 981 N; contact4 lifts
    x = np.zeros(prob.n)
    x[0:3] = [0.0, 0.0, 1.0]
    for k, nm in enumerate(NAMES):
        x[3 + 9 * k:6 + 9 * k] = [0.0, 0.0, Fz[nm]]
        x[6 + 9 * k:9 + 9 * k] = pts[nm]
        x[9 + 9 * k:12 + 9 * k] = [0.0, 0.0, 1.0]
    xd = torch.from_numpy(np.tile(x, (64, 1))).to(cuda_device)
    g = prob.EvaluateConstraints(xd).cpu().numpy()[0]
    # statics rows: sum F + m g = 0; the moment rows equal the torque of the contact forces about the CoM
    assert g[0] == 0.0 and g[1] == 0.0 and abs(g[2]) <= 1e-12
    tau = sum(np.cross(np.array(pts[nm]) - x[0:3], [0, 0, x[3 + 9 * k + 2]]) for k, nm in enumerate(NAMES))
    assert np.allclose(g[3:6], tau, atol=1e-12)
    for j, nm in enumerate(sorted(NAMES)):
        k = NAMES.index(nm)
        fz = x[3 + 9 * k + 2]
        thr = 0.0 if nm == "contact4" else 20.0
        assert g[6 + 2 * j] == -fz + thr                      # -F.n + F_thr
        assert g[6 + 2 * j + 1] == 0.0 - 0.5 * fz             # |F_t| - mu F.n
        assert g[6 + 2 * j] <= 0.0 and g[6 + 2 * j + 1] <= 0.0

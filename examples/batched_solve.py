#!/usr/bin/env python
"""The reference's examples/python flow (build a planner, set bounds and weights, Solve(), print the solution) for a BATCH of
instances on the GPU.  IPOPT is not part of this repository: Solve() runs the lock-step interior-point stand-in of
centroidalplanner_b200/lockstep_solver.py over the batched evaluator.

    python examples/batched_solve.py            # needs a CUDA device
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import centroidalplanner_b200 as cpl
from centroidalplanner_b200.lockstep_solver import STATUS_NAMES, default_start

contacts = ["contact1", "contact2", "contact3", "contact4"]
ground = cpl.Ground()
ground.SetGroundZ(0.1)
ground.SetMu(0.5)
planner = cpl.BatchedCentroidalPlanner(contacts, 100.0, ground)       # tests/TestBasic.cpp:64-99
planner.SetCoMWeight(2.0)
planner.SetForceWeight(0.0)
for c in contacts:
    planner.SetPosBounds(c, [-0.3, -0.3, 0.0], [0.3, 0.3, 1.0])
planner.SetManipulationWrench([100.0, 0, 0, 0, 0, 100.0])

N = 256                                                                 # 256 starting points, one lock-step solve
x0 = default_start(planner.GetCplProblem(), N, device="cuda")
x0[1:] += 0.05 * torch.randn_like(x0[1:])
sols = planner.Solve(x0)
rep = planner.last_solve
print(f"{N} instances, {rep.rounds} lock-step rounds, {rep.evaluations} kernel launches; "
      f"status: { {STATUS_NAMES[int(k)]: int(v) for k, v in zip(*np.unique(rep.status.cpu().numpy(), return_counts=True))} }")
print(planner.GetCplProblem().FormatSolution(sols[0]))                  # the text `std::cout << sol` prints in the reference
F_sum = sum(v["force_value"] for v in sols[0]["contact_values_map"].values())
print("sum of contact forces:", F_sum, "(manipulation wrench + weight = [100, 0, 981])")

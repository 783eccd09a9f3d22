#!/usr/bin/env python
"""Counterpart of the reference's examples/python/test.py for the batched evaluator: the same four-contact problem on
flat ground and on a superquadric, evaluated for a batch of random points on the GPU (IPOPT would sit on top).

    python examples/batched_eval.py            # needs a CUDA device
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import centroidalplanner_b200 as cpl
from centroidalplanner_b200 import synthetic

contacts = ["contact1", "contact2", "contact3", "contact4"]

ground = cpl.Ground()
ground.SetGroundZ(0.1)
ground.SetMu(0.5)
prob = cpl.BatchedCplProblem(contacts, 100.0, ground)
prob.SetCoMWeight(2.0)
prob.SetForceWeight(0.0)
prob.SetManipulationWrench([100.0, 0, 0, 0, 0, 100.0])
for c in contacts:
    prob.SetPosBounds(c, [-0.3, -0.3, 0.0], [0.3, 0.3, 1.0])

N = 8192
x = torch.from_numpy(synthetic.ground_batch(N)).cuda()            # (N, 39): CoM, then F, p, n per contact
out = prob.eval(x, g=True, jac=True, cost=True, grad=True)        # one fused kernel launch
torch.cuda.synchronize()
iRow, jCol = prob.GetJacobianStructure()                          # what IPOPT's eval_jac_g(values=NULL) would receive
print(f"n={prob.n} m={prob.m} nnz={prob.nnz}; first triplets: {list(zip(iRow[:4], jCol[:4]))}")
print("instance 0: force balance residual", out["g"][0, :3].cpu().numpy(), " cost", float(out["cost"][0]))
print("solution layout of instance 0:", {k: v for k, v in prob.GetSolution(x[0].cpu().numpy())["contact_values_map"]["contact1"].items()})

sq = cpl.Superquadric()
sq.SetParameters([0.0, 0.0, 1.0], [0.3, 0.3, 10.0], [10.0, 10.0, 10.0])
sq.SetMu(0.5)
prob2 = cpl.BatchedCplProblem(contacts, 100.0, sq)
x2 = torch.from_numpy(synthetic.superquadric_batch(N)).cuda()
g2 = prob2.EvaluateConstraints(x2)
print("superquadric: surface-distance rows of instance 0:", g2[0, 6::6].cpu().numpy())

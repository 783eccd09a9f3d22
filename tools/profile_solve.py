"""Where a lock-step solve spends its time on the GPU: torch.profiler table of one 4,096-instance solve (tools/solve_batch.py's
gpu arm).  Diagnostic only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from centroidalplanner_b200.lockstep_solver import LockStepInteriorPoint  # noqa: E402
import test_solve as ts  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prob, names, par = ts.product_problem("ground")
x0 = ts.starts(prob, N, seed=2025, device=torch.device("cuda:0"))
solver = LockStepInteriorPoint()
solver.Solve(prob, x0[:64])
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    res = solver.Solve(prob, x0)
    torch.cuda.synchronize()
print("rounds", res.rounds, "evaluations", res.evaluations)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="cpu_time_total", row_limit=15, max_name_column_width=60))

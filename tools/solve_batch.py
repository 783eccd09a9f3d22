"""BASELINE.json configs[4] stand-in: end-to-end solves/s for 4,096 instances through the lock-step solve driver.

NOT an IPOPT measurement -- IPOPT is absent from the image.  What runs is
`centroidalplanner_b200.lockstep_solver.LockStepInteriorPoint` (the interior-point scheme IPOPT implements, batched)
on the TestBasic ground problem (configs[0] parameters) from N different starting points:

  gpu arm  : evaluator = the CUDA path (cplb_eval_device), linear algebra = torch on the same device, x never leaves HBM;
  cpu arm  : the identical driver on the host -- evaluator = the C oracle on all host threads, linear algebra = torch CPU --
             on a bounded sample of the same starting points.

Prints one JSON line.  `python tools/solve_batch.py [--instances 4096] [--cpu-sample 256] [--case ground|superquadric|com_planner]`
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from centroidalplanner_b200.lockstep_solver import SUCCESS, LockStepInteriorPoint  # noqa: E402
import test_solve as ts  # noqa: E402  (the TestBasic setups)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=4096)
    ap.add_argument("--cpu-sample", type=int, default=256)
    ap.add_argument("--case", default="ground", choices=list(ts.SETUPS))
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--no-gpu", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--solver", default="native", choices=["native", "torch"],
                    help="native: cplb_solve_device (csrc/cplb_solver.cu); torch: the previous driver (lockstep_solver.py)")
    ap.add_argument("--tail", type=int, default=-1,
                    help="native solver: working-set size at which the tail takes over (-1: what the GPU holds at once; 0: never)")
    a = ap.parse_args()
    out = {"metric": "end-to-end solves/s (lock-step interior-point stand-in, NOT IPOPT)", "case": a.case, "instances": a.instances}
    # one process per GPU under torchrun: instances shard by index (no collective on the solve path), time = max over ranks
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    op, names, par = ts.oracle_problem(a.case)
    op.nthreads = len(os.sched_getaffinity(0))
    x0 = ts.starts(op, a.instances, seed=2025)
    solver = LockStepInteriorPoint()
    out["solver"] = a.solver

    if not a.no_gpu:
        if a.solver == "native":
            import centroidalplanner_b200 as cpl
            solver = cpl.NativeInteriorPoint(tail_instances=a.tail)
        from centroidalplanner_b200 import sharding
        prob, _, _ = ts.product_problem(a.case)
        dev = torch.device("cuda", local)
        lo, hi = sharding.local_range(a.instances, rank, world)
        xg = x0[lo:hi].to(dev)
        solver.Solve(prob, xg[:64])          # warm-up: cuSOLVER/cuBLAS handles, kernels
        best = None
        for _ in range(a.repeats):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            l0 = prob.launch_count()
            t = time.perf_counter()
            res = solver.Solve(prob, xg)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t, float(res.rounds)], dtype=torch.float64, device=dev)
            if world > 1:
                every = [torch.zeros_like(dt) for _ in range(world)]
                dist.all_gather(every, dt)
                per_rank = [[round(float(e[0]), 4), int(e[1])] for e in every]      # (seconds, rounds) of every rank
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            else:
                per_rank = [[round(float(dt[0]), 4), int(dt[1])]]
            dt = float(dt[0].item())
            best = dt if best is None else min(best, dt)
        ok = torch.tensor([int((res.status == SUCCESS).sum())], device=dev)
        if world > 1:
            dist.all_reduce(ok)
        ok = int(ok.item())
        ts.check_expectations(a.case, names, par, res.x[res.status == SUCCESS].cpu().numpy())
        out["gpu"] = {"n_gpus": world, "solves_per_s": a.instances / best, "seconds": best, "succeeded": ok, "rounds": res.rounds,
                      "kernel_launches": prob.launch_count() - l0, "instance_evaluations": res.instance_evaluations,
                      "iterations_median": float(res.iterations.double().median()), "iterations_max": int(res.iterations.max()),
                      "per_rank_seconds_rounds": per_rank, "tail_instances": getattr(res, "tail_instances", 0),
                      "max_constr_viol": float(res.constr_viol[res.status == SUCCESS].max())}

    if rank != 0:
        return
    if a.no_cpu:
        print(json.dumps(out))
        return
    solver = LockStepInteriorPoint()
    n_cpu = min(a.cpu_sample, a.instances)
    t = time.perf_counter()
    rc = solver.Solve(op, x0[:n_cpu])
    dt = time.perf_counter() - t
    out["cpu"] = {"solves_per_s": n_cpu / dt, "seconds": dt, "sample": n_cpu, "succeeded": int((rc.status == SUCCESS).sum()),
                  "threads": op.nthreads, "torch_threads": torch.get_num_threads(), "rounds": rc.rounds,
                  "kind": "same driver, C oracle evaluator + torch CPU linear algebra"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

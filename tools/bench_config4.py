#!/usr/bin/env python
"""BASELINE.json configs[3]: 1,048,576 synthetic 8-contact (hands+feet) flat-ground instances sharded by index across
the GPUs of one box (strong scaling: the total is fixed, rank r evaluates sharding.local_range(N, r, G)).

    python tools/bench_config4.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/bench_config4.py

No collective on the evaluation path; timing = CUDA events on each rank's stream, max over ranks (one all_reduce after
the timed region).  Contact names are given "r_..." first so that vector order != sorted order."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import centroidalplanner_b200 as cpl  # noqa: E402
from centroidalplanner_b200 import sharding, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--layout", default="component", choices=["component", "instance"])
    ap.add_argument("--no-inputs-ready", action="store_true", help="plain stream order (no CPLB_DEVICE_INPUTS_READY)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    env = cpl.Ground()
    env.SetGroundZ(0.1)
    env.SetMu(0.5)
    prob = cpl.BatchedCplProblem(synthetic.NAMES8, 100.0, env, device=local)
    prob.SetManipulationWrench(synthetic.TESTBASIC["wrench"])
    lo, hi = sharding.local_range(a.instances, rank, world)
    n_loc = hi - lo
    # every rank generates only its own shard (seeded per 65,536-instance block so the batch does not depend on G)
    blocks = [synthetic.ground_batch(65536, 8, seed=1004 + b) for b in range(lo // 65536, (hi + 65535) // 65536)]
    x = np.concatenate(blocks)[lo - (lo // 65536) * 65536:][:n_loc]
    layout = cpl.COMPONENT_MAJOR if a.layout == "component" else cpl.INSTANCE_MAJOR
    shp = (lambda L: (L, n_loc)) if layout == cpl.COMPONENT_MAJOR else (lambda L: (n_loc, L))
    per = 8 * (prob.n + prob.m + prob.nnz)
    sets = max(2, int(np.ceil(4 * 126 * 2**20 / (per * n_loc))))
    xd = torch.from_numpy(np.ascontiguousarray(x.T) if layout == cpl.COMPONENT_MAJOR else x).to(dev)
    xs = [xd.clone() for _ in range(sets)]
    outs = [{"g": torch.empty(shp(prob.m), dtype=torch.float64, device=dev),
             "jac": torch.empty(shp(prob.nnz), dtype=torch.float64, device=dev)} for _ in range(sets)]
    for i in range(a.warmup):
        prob.eval(xs[i % sets], g=True, jac=True, layout=layout, out=outs[i % sets], inputs_ready=not a.no_inputs_ready)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        s = (a.warmup + i) % sets
        prob.eval(xs[s], g=True, jac=True, layout=layout, out=outs[s], inputs_ready=not a.no_inputs_ready)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        gbs = per * a.instances / ms / 1e6
        print(json.dumps({"workload": "configs[3]: 1,048,576 x 8-contact flat ground, sharded by index (strong scaling)",
                          "n_gpus": world, "instances": a.instances, "instances_per_gpu": n_loc, "layout": a.layout,
                          "ms_per_step": ms, "value": a.instances / (ms * 1e-3), "unit": "instances/s",
                          "algorithmic_GBs_total": gbs, "frac_of_n_gpus_x_measured_peak": gbs / (world * peak), "buffer_sets": sets,
                          "inputs_ready": not a.no_inputs_ready}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

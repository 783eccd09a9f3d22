#!/usr/bin/env python
"""Small driver for profiling: a few launches of one configuration through the C ABI (device path).

    python tools/run_eval.py --case ground4 --layout instance --n 65536 --steps 5
Used under `ncu` (one kernel family per run) and for quick A/B timing with CUDA events."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import centroidalplanner_b200 as cpl  # noqa: E402
from helpers import make_pair  # noqa: E402  (only for the shared parameter sets; the oracle side is unused here)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="ground4")
    ap.add_argument("--layout", default="instance", choices=["instance", "component"])
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--sets", type=int, default=0, help="buffer sets to rotate through (0 = enough to exceed L2 8x)")
    ap.add_argument("--all", action="store_true", help="also cost + gradient")
    ap.add_argument("--ready", action="store_true", help="pass inputs_ready=True (x is not produced by the preceding kernel)")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    prob, _, gen = make_pair(a.case, rich=False)
    x = gen(min(a.n, 1 << 16))
    if a.n > x.shape[0]:
        x = np.tile(x, ((a.n + x.shape[0] - 1) // x.shape[0], 1))[: a.n]
    layout = cpl.INSTANCE_MAJOR if a.layout == "instance" else cpl.COMPONENT_MAJOR
    per = 8 * (prob.n + prob.m + prob.nnz) * a.n
    sets = a.sets or max(2, int(np.ceil(8 * 126 * 2**20 / per)))
    xd = torch.from_numpy(x).cuda()
    if layout == cpl.COMPONENT_MAJOR:
        xd = xd.t().contiguous()
    xs = [xd.clone() for _ in range(sets)]
    shp = (lambda L: (a.n, L)) if layout == cpl.INSTANCE_MAJOR else (lambda L: (L, a.n))
    outs = [{"g": torch.empty(shp(prob.m), dtype=torch.float64, device="cuda"),
             "jac": torch.empty(shp(prob.nnz), dtype=torch.float64, device="cuda")} for _ in range(sets)]
    if a.all:
        for o in outs:
            o["cost"] = torch.empty(a.n, dtype=torch.float64, device="cuda")
            o["grad"] = torch.empty(shp(prob.n), dtype=torch.float64, device="cuda")
    for i in range(a.warmup):
        prob.eval(xs[i % sets], g=True, jac=True, cost=a.all, grad=a.all, layout=layout, out=outs[i % sets])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        s = (a.warmup + i) % sets
        prob.eval(xs[s], g=True, jac=True, cost=a.all, grad=a.all, layout=layout, out=outs[s], inputs_ready=a.ready)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    alg = per + (8 * (prob.n + 1) * a.n if a.all else 0)
    print(f"{a.case} {a.layout} N={a.n} steps={a.steps} sets={sets}: {ms*1e3:.2f} us/step, "
          f"{a.n/ms/1e3:.1f} M inst/s, {alg/ms/1e6:.0f} GB/s algorithmic ({alg/ms/1e6/6449.7*100:.1f}% of 6449.7)")


if __name__ == "__main__":
    main()

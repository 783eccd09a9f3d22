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
    ap.add_argument("--im-kernel", default="auto", choices=["auto", "warp", "cta"], help="cplb_set_instance_major_kernel")
    ap.add_argument("--cm-kernel", default="auto", choices=["auto", "split", "whole"], help="cplb_set_component_major_kernel")
    ap.add_argument("--graph", action="store_true", help="time CUDA-graph replays of one rotation over the buffer sets (no Python launch path in the step)")
    ap.add_argument("--slices", default="full", choices=["full", "packed", "computed"],
                    help="instance-major Jacobian slice format: all structural slots, CPLB_JAC_PACKED, CPLB_JAC_COMPUTED")
    ap.add_argument("--perinst", action="store_true", help="per-instance constraint parameters (mass, wrench, mu, thresholds, ground z)")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    prob, _, gen = make_pair(a.case, rich=False)
    prob.SetInstanceMajorKernel(a.im_kernel)
    prob.SetComponentMajorKernel(a.cm_kernel)
    x = gen(min(a.n, 1 << 16))
    if a.n > x.shape[0]:
        x = np.tile(x, ((a.n + x.shape[0] - 1) // x.shape[0], 1))[: a.n]
    layout = cpl.INSTANCE_MAJOR if a.layout == "instance" else cpl.COMPONENT_MAJOR
    jp = {"full": False, "packed": True, "computed": "computed"}[a.slices]
    jac_len = prob._jac_len(jp)
    per = 8 * (prob.n + prob.m + jac_len) * a.n   # bytes the launch has to move (x in, g and the Jacobian slices out)
    sets = a.sets or max(2, int(np.ceil(8 * 126 * 2**20 / per)))
    xd = torch.from_numpy(x).cuda()
    if layout == cpl.COMPONENT_MAJOR:
        xd = xd.t().contiguous()
    xs = [xd.clone() for _ in range(sets)]
    shp = (lambda L: (a.n, L)) if layout == cpl.INSTANCE_MAJOR else (lambda L: (L, a.n))
    outs = [{"g": torch.empty(shp(prob.m), dtype=torch.float64, device="cuda"),
             "jac": torch.empty(shp(jac_len), dtype=torch.float64, device="cuda")} for _ in range(sets)]
    if a.all:
        for o in outs:
            o["cost"] = torch.empty(a.n, dtype=torch.float64, device="cuda")
            o["grad"] = torch.empty(shp(prob.n), dtype=torch.float64, device="cuda")
    pi = None
    if a.perinst:
        nc = (prob.n - 3) // 9
        rng = np.random.default_rng(1)
        shp_p = (lambda L: (a.n, L)) if layout == cpl.INSTANCE_MAJOR else (lambda L: (L, a.n))
        pi = {"mass": rng.uniform(20, 150, a.n), "wrench": rng.uniform(-50, 50, shp_p(6)), "mu": rng.uniform(0.2, 1.2, a.n),
              "force_threshold": rng.uniform(0, 30, shp_p(nc))}
        if a.case.startswith("ground"):
            pi["ground_z"] = rng.uniform(-0.2, 0.4, a.n)
        pi = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in pi.items()}
    for i in range(a.warmup):
        prob.eval(xs[i % sets], g=True, jac=True, cost=a.all, grad=a.all, layout=layout, out=outs[i % sets], per_instance=pi, jac_packed=jp)
    torch.cuda.synchronize()
    def step(i):
        s = (a.warmup + i) % sets
        prob.eval(xs[s], g=True, jac=True, cost=a.all, grad=a.all, layout=layout, out=outs[s], inputs_ready=a.ready, per_instance=pi, jac_packed=jp)

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if a.graph:
        stream = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(stream)
        with torch.cuda.stream(side):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                for i in range(sets):
                    step(i)
        stream.wait_stream(side)
        graph.replay()
        torch.cuda.synchronize()
        reps = max(1, a.steps // sets)
        a.steps = reps * sets
        e0.record()
        for _ in range(reps):
            graph.replay()
        e1.record()
    else:
        e0.record()
        for i in range(a.steps):
            step(i)
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    alg = per + (8 * (prob.n + 1) * a.n if a.all else 0)
    if a.graph:
        a.case += " [graph]"
    tag = f" im={a.im_kernel}" if a.layout == "instance" else f" cm={a.cm_kernel}"
    print(f"{a.case} {a.layout}{tag}{' perinst' if a.perinst else ''}{' all4' if a.all else ''}{'' if a.slices == 'full' else ' ' + a.slices} N={a.n} steps={a.steps} sets={sets}: {ms*1e3:.2f} us/step, "
          f"{a.n/ms/1e3:.1f} M inst/s, {alg/ms/1e6:.0f} GB/s algorithmic ({alg/ms/1e6/6449.7*100:.1f}% of 6449.7)")


if __name__ == "__main__":
    main()

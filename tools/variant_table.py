#!/usr/bin/env python
"""Every kernel variant against its own HBM roofline: problem shape x environment x layout x batch size x outputs x
parameter mode, device path, CUDA events over steps that rotate through buffer sets larger than L2.

    python tools/variant_table.py > profiles/r01_variants.md

Algorithmic bytes per instance: 8 (n + m + nnz) for g + Jacobian, + 8 (n + 1) with cost + gradient, and in
per-instance parameter mode + 8 (8 + nc + E) (mass, wrench, mu, F_thr[nc], Ground z) and, with the cost, + 8 (8 nc + 4)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import json  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

import centroidalplanner_b200 as cpl  # noqa: E402
from helpers import make_pair  # noqa: E402  (shared parameter sets only; the oracle side is unused)

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6449.7


def per_instance_arrays(prob, N, nc, env, want_cost, layout, rng):
    """Random but valid per-instance parameters, laid out like x."""
    def arr(L, lo, hi):
        a = rng.uniform(lo, hi, (N, L)) if L > 1 else rng.uniform(lo, hi, (N,))
        t = torch.from_numpy(a).cuda()
        return t.t().contiguous() if (layout == cpl.COMPONENT_MAJOR and L > 1) else t
    d = {"mass": arr(1, 50, 150), "wrench": arr(6, -100, 100), "mu": arr(1, 0.3, 0.9), "force_threshold": arr(nc, 0, 30)}
    nbytes = 8 * (8 + nc)
    if env == "ground":
        d["ground_z"] = arr(1, -0.1, 0.2)
        nbytes += 8
    if want_cost:
        d.update({"com_ref": arr(3, -0.1, 0.1), "com_weight": arr(1, 0.5, 2), "pos_ref": arr(3 * nc, -0.3, 0.3),
                  "force_ref": arr(3 * nc, -50, 300), "pos_weight": arr(nc, 0.5, 2), "force_weight": arr(nc, 0, 0.01)})
        nbytes += 8 * (8 * nc + 4)
    return d, nbytes


def measure(case, layout, N, want_all, per_inst, steps=40, warmup=5):
    prob, o, gen = make_pair(case, rich=False)
    nc = (prob.n - 3) // 9
    env = {cpl.ENV_NONE: "none", cpl.ENV_GROUND: "ground", cpl.ENV_SUPERQUADRIC: "superquadric"}[prob._env_kind]
    x = gen(min(N, 1 << 16))
    if N > x.shape[0]:
        x = np.tile(x, ((N + x.shape[0] - 1) // x.shape[0], 1))[:N]
    per = 8 * (prob.n + prob.m + prob.nnz) + (8 * (prob.n + 1) if want_all else 0)
    rng = np.random.default_rng(3)
    pi, extra = (per_instance_arrays(prob, N, nc, env, want_all, layout, rng) if per_inst else (None, 0))
    per += extra
    sets = max(3, int(np.ceil(4 * 126 * 2**20 / (per * N))))
    xd = torch.from_numpy(x).cuda()
    if layout == cpl.COMPONENT_MAJOR:
        xd = xd.t().contiguous()
    xs = [xd.clone() for _ in range(sets)]
    shp = (lambda L: (N, L)) if layout == cpl.INSTANCE_MAJOR else (lambda L: (L, N))
    outs = []
    for _ in range(sets):
        d = {"g": torch.empty(shp(prob.m), dtype=torch.float64, device="cuda"), "jac": torch.empty(shp(prob.nnz), dtype=torch.float64, device="cuda")}
        if want_all:
            d["cost"] = torch.empty(N, dtype=torch.float64, device="cuda")
            d["grad"] = torch.empty(shp(prob.n), dtype=torch.float64, device="cuda")
        outs.append(d)
    kw = dict(g=True, jac=True, cost=want_all, grad=want_all, layout=layout, per_instance=pi, inputs_ready=True)
    g = torch.cuda.CUDAGraph()                     # replay a graph of the rotation so the Python call path is not timed
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(warmup):
            prob.eval(xs[i % sets], out=outs[i % sets], **kw)
        s.synchronize()
        with torch.cuda.graph(g, stream=s):
            for i in range(sets):
                prob.eval(xs[i], out=outs[i], **kw)
        reps = max(1, steps // sets)
        g.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            g.replay()
        e1.record(s)
        s.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * sets)
    gbs = per * N / us / 1e3
    return nc, env, per, us, gbs


def main():
    torch.cuda.set_device(0)
    print(f"# Kernel variants vs the HBM roofline ({PEAK:.1f} GB/s measured copy bandwidth)\n")
    print("Device path, CUDA-graph replay of a rotation through buffer sets > 4x L2, `inputs_ready` (a queue of independent batches),"
          " CUDA events on the launching stream. `B/inst` = algorithmic bytes per instance (SURVEY §8(d)).\n")
    print("| problem | contacts | layout | N | outputs | parameters | B/inst | us/launch | M inst/s | GB/s | of roofline |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    cases = ["ground4", "superquadric4", "noenv4", "ground8", "superquadric8", "noenv8", "ground1", "noenv2", "ground12", "superquadric32"]
    for case in cases:
        for layout, lname in ((cpl.COMPONENT_MAJOR, "component-major"), (cpl.INSTANCE_MAJOR, "instance-major")):
            for N in (65536, 1 << 20):
                for want_all in (False, True):
                    for per_inst in (False, True):
                        if per_inst and case.startswith("superquadric"):
                            pass                               # superquadric shape stays shared; the other parameters vary
                        if (want_all or per_inst) and case not in ("ground4", "superquadric4", "noenv4", "ground8"):
                            continue
                        if case == "superquadric32" and N > 65536:
                            continue
                        try:
                            nc, env, per, us, gbs = measure(case, layout, N, want_all, per_inst)
                        except Exception as e:                 # noqa: BLE001
                            print(f"| {case} | | {lname} | {N} | | | | error: {str(e)[:60]} | | | |")
                            continue
                        print(f"| {env} | {nc} | {lname} | {N:,} | {'g+J+cost+grad' if want_all else 'g+J'} | {'per-instance' if per_inst else 'shared'} | "
                              f"{per:,} | {us:.1f} | {N / us:.0f} | {gbs:.0f} | {100 * gbs / PEAK:.1f} % |", flush=True)
                        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

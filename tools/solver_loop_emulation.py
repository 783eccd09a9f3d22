#!/usr/bin/env python
"""EMULATION of BASELINE.json configs[4] ("end-to-end IPOPT solves/sec for 4,096 instances") -- NOT a measurement of it.

IPOPT is not installed in this image, so no solve can run.  What this script times is the part of a solve that this
repository replaces: per solver iteration IPOPT issues eval_f, eval_grad_f, eval_g, eval_jac_g on the current x; here a
fixed schedule of R such callback rounds runs for 4,096 instances with a stand-in x update (a damped gradient step --
it is NOT an interior-point step and converges to nothing in particular), once through the batched GPU evaluator
(cplb_eval_host: host buffers in and out, what lock-step solver threads would share) and once through the CPU path
(the reference's own sources, oracle/_ref, all host cores; falls back to the oracle port).  Output: callback rounds/s.
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=4096)
    ap.add_argument("--rounds", type=int, default=50)
    ap.add_argument("--case", default="ground4")
    a = ap.parse_args()
    from helpers import CASES, OracleProblem, configure, make_pair
    from oracle import cpl_ref_py

    prob, o, gen = make_pair(a.case, rich=False)
    x0 = gen(a.instances)
    cores = len(os.sched_getaffinity(0))

    def run(evaluate):
        x = x0.copy()
        t0 = time.perf_counter()
        for _ in range(a.rounds):
            out = evaluate(x)                      # eval_f + eval_grad_f + eval_g + eval_jac_g of every instance
            x = x - 1e-4 * out["grad"]            # stand-in for the solver's step
        return (time.perf_counter() - t0) / a.rounds, x

    # pinned host buffers allocated once, as cplb::solver::InstanceBatch does for the solver threads
    import ctypes as C

    from centroidalplanner_b200 import _cabi

    lib = _cabi.load()

    def pinned(shape):
        ptr = C.c_void_p()
        assert lib.cplb_host_alloc(int(np.prod(shape)) * 8, C.byref(ptr)) == 0
        return np.ctypeslib.as_array((C.c_double * int(np.prod(shape))).from_address(ptr.value)).reshape(shape)

    N = a.instances
    hx = pinned((N, prob.n))
    bufs = {"g": pinned((N, prob.m)), "jac": pinned((N, prob.nnz)), "cost": pinned((N,)), "grad": pinned((N, prob.n))}

    def gpu_eval(x):
        hx[...] = x
        return prob.eval(hx, g=True, jac=True, cost=True, grad=True, out=bufs)

    gpu_eval(x0)  # first call creates the streams and the device staging buffers
    gpu_s, xg = run(gpu_eval)
    if cpl_ref_py.available():
        names, env_name = CASES[a.case][0], CASES[a.case][1]
        rp = cpl_ref_py.RefProblem(names, env_name, 100.0)
        configure(rp, rp if env_name != "none" else None, names, env_name, False)
        cpu_kind, cpu_eval = "reference sources (oracle/_ref)", (lambda x: rp.eval_batch(x, nthreads=cores))
    else:
        cpu_kind, cpu_eval = "oracle port", (lambda x: o.eval_batch(x, nthreads=cores))
    cpu_s, xc = run(cpu_eval)
    print(json.dumps({
        "what": "EMULATED solver-callback rounds (not IPOPT solves; IPOPT is absent from the image)",
        "instances": a.instances, "rounds": a.rounds, "case": a.case,
        "gpu_batched": {"ms_per_round": 1e3 * gpu_s, "instance_rounds_per_s": a.instances / gpu_s},
        "cpu": {"kind": cpu_kind, "cores": cores, "ms_per_round": 1e3 * cpu_s, "instance_rounds_per_s": a.instances / cpu_s},
        "same_trajectory_bits": bool(np.array_equal(xg, xc, equal_nan=True)),
    }))


if __name__ == "__main__":
    main()

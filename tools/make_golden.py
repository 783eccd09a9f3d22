#!/usr/bin/env python
"""Generate tests/golden/ref_vectors.npz from oracle/_ref/libcpl_ref.so -- the reference's own sources
(/root/reference/src, compiled in place against oracle/refshim) driven through ifopt-style
Problem::EvaluateConstraints / EvalNonzerosOfJacobian / EvaluateCostFunction[Gradient].

Run in the build container (where /root/reference exists):  make -C oracle ref && python tools/make_golden.py
The fixture travels with the repo; neither the tests nor the GPU box need /root/reference."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

from helpers import CASES, configure  # noqa: E402
from oracle import cpl_ref_py  # noqa: E402

N = 24


def main():
    assert cpl_ref_py.available(), "build oracle/_ref first: make -C oracle ref"
    out = {}
    for case in sorted(CASES):
        names, env_name, gen = CASES[case][:3]
        extra = CASES[case][3] if len(CASES[case]) > 3 else None
        rp = cpl_ref_py.RefProblem(names, env_name, 100.0)
        configure(rp, rp if env_name != "none" else None, names, env_name, True, extra)
        x = gen(N)
        x[-1] = 0.0  # the reference's default start point (Variable3D.cpp:8-10): NaN pattern of FrictionCone
        r = rp.eval_batch(x)
        iRow, jCol = rp.structure()
        xl, xu, gl, gu = rp.bounds()
        for k, v in dict(x=x, g=r["g"], jac=r["jac"], cost=r["cost"], grad=r["grad"], iRow=iRow, jCol=jCol, xl=xl, xu=xu,
                         gl=gl, gu=gu).items():
            out[f"{case}/{k}"] = v
    dst = os.path.join(ROOT, "tests", "golden", "ref_vectors.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes,", len(CASES), "cases x", N, "instances")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""How much of the expanded-square allowance (tests/helpers.py::expanded_square_bound) the CUDA path needs.

Contacts are placed at relative distances 1e-1 .. 1e-7 from the superquadric's centre planes (where the reference's diagonal
normal-Jacobian entries are dominated by the rounding of C^2 + p^2 - 2 C p, Superquadric.cpp:99-100,153-154,207-208).  For every
distance band: entries beyond the plain 1e-12 bar, share of entries bit-identical to the oracle, and the largest
|gpu - oracle| / (eps * amplification * |exact|) -- the factor the allowance must carry.  Run on the GPU box."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import helpers  # noqa: E402
from helpers import make_pair, pow_downstream_masks  # noqa: E402


def main():
    torch.cuda.set_device(0)
    prob, o, gen = make_pair("superquadric4", rich=False)
    C = np.array([0.0, 0.0, 1.0])
    N = 20000
    rng = np.random.default_rng(3)
    _, jm = pow_downstream_masks(o)
    print("| |p_z - C_z| band | entries beyond 1e-12 | bit-identical entries | max err / (eps amp |exact|) |")
    print("|---|---|---|---|")
    worst = 0.0
    for k in range(1, 8):
        x = gen(N)
        lo, hi = 10.0 ** (-k), 10.0 ** (-k + 1)
        for c in range(4):
            sgn = rng.choice([-1.0, 1.0], N)
            x[:, 3 + 9 * c + 5] = C[2] + sgn * rng.uniform(lo, hi, N)
        want = o.eval_batch(x, want=("jac",), nthreads=8)["jac"]
        got = prob.eval(torch.from_numpy(x).cuda(), g=False, jac=True)["jac"].cpu().numpy()
        helpers.Q5_FACTOR = 1.0
        unit = helpers.expanded_square_bound(o, x)            # eps * amp * |exact| per diagonal entry, 0 elsewhere
        err = np.abs(got - want)
        diag = unit > 0
        beyond = int((err[:, jm] > 1e-12 * np.abs(want[:, jm])).sum())
        same = float((got[:, jm].view(np.int64) == want[:, jm].view(np.int64)).mean())
        with np.errstate(all="ignore"):
            ratio = np.where(diag & np.isfinite(want), err / unit, 0.0)
        off = err[:, jm & ~diag.any(axis=0)]
        off_rel = float((off / np.maximum(np.abs(want[:, jm & ~diag.any(axis=0)]), 1e-300)).max())
        worst = max(worst, float(ratio.max()))
        print(f"| [{lo:.0e}, {hi:.0e}) | {beyond} of {int(jm.sum()) * N} | {100 * same:.1f} % | {float(ratio.max()):.2f} (other pow-downstream entries: {off_rel:.1e} relative) |")
    print(f"\nlargest factor needed: {worst:.2f}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Largest deviation of the GPU evaluation from the CPU oracle over ALL instances of the Superquadric benchmark batches
(config 3: 65,536 x 4 contacts; and the 8-contact variant).  Outputs without a pow() upstream must be bit-identical; for
the others the script prints the maximum and median relative deviation.  Recorded in DESIGN.md section 5."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import centroidalplanner_b200 as cpl
from helpers import make_pair, pow_downstream_masks, same_bits
for case in ("superquadric4","superquadric8"):
    prob,o,gen=make_pair(case, rich=False)
    N=65536; x=gen(N)
    o.set_call_all_pairs(0)
    want=o.eval_batch(x, want=("g","jac"), nthreads=16)
    xd=torch.from_numpy(x).cuda()
    for layout in (cpl.COMPONENT_MAJOR, cpl.INSTANCE_MAJOR):
        xin = xd.t().contiguous() if layout==cpl.COMPONENT_MAJOR else xd
        out=prob.eval(xin,g=True,jac=True,layout=layout); torch.cuda.synchronize()
        g=out["g"].cpu().numpy(); j=out["jac"].cpu().numpy()
        if layout==cpl.COMPONENT_MAJOR: g=g.T; j=j.T
        gm,jm=pow_downstream_masks(o)
        eg=np.abs(g[:,gm]-want["g"][:,gm])/np.maximum(np.abs(want["g"][:,gm]),1.0)
        ej=np.abs(j[:,jm]-want["jac"][:,jm])/np.abs(want["jac"][:,jm])
        print(case, "layout",layout,"pow-free bit-identical:", same_bits(g[:,~gm],want["g"][:,~gm]) and same_bits(j[:,~jm],want["jac"][:,~jm]),
              " max rel dev g(pow rows, floor 1):", eg.max(), " jac(pow slots):", np.nanmax(ej), " median:", np.nanmedian(ej))

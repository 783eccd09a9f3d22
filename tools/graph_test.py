"""Diagnostic: host issue rate of back-to-back evaluations from Python vs the device time per step, and the same steps
replayed from a CUDA graph (what bench.py times).  python tools/graph_test.py  (needs a GPU)"""
import sys, time; import os; ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
import numpy as np, torch
import centroidalplanner_b200 as cpl
from helpers import make_pair
torch.cuda.set_device(0)
prob,_,gen = make_pair("ground4", rich=False)
N=65536; x=gen(N)
for layout in (cpl.COMPONENT_MAJOR, cpl.INSTANCE_MAJOR):
    sets=9
    xd=torch.from_numpy(x).cuda()
    if layout==cpl.COMPONENT_MAJOR: xd=xd.t().contiguous()
    shp=(lambda L:(N,L)) if layout==cpl.INSTANCE_MAJOR else (lambda L:(L,N))
    xs=[xd.clone() for _ in range(sets)]
    outs=[{"g":torch.empty(shp(prob.m),dtype=torch.float64,device='cuda'),"jac":torch.empty(shp(prob.nnz),dtype=torch.float64,device='cuda')} for _ in range(sets)]
    def step(i): prob.eval(xs[i%sets],g=True,jac=True,layout=layout,out=outs[i%sets])
    for i in range(20): step(i)
    torch.cuda.synchronize()
    K=900
    t0=time.perf_counter()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K): step(i)
    e1.record()
    t1=time.perf_counter()
    torch.cuda.synchronize()
    print(f"layout {layout}: host issue {1e6*(t1-t0)/K:.2f} us/launch, device {1e3*e0.elapsed_time(e1)/K:.2f} us/step")
    # graph
    g=torch.cuda.CUDAGraph()
    s=torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(3): step(i)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for i in range(sets): step(i)
    torch.cuda.synchronize()
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(K//sets): g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"layout {layout}: graph replay {1e3*e0.elapsed_time(e1)/(K//sets*sets):.2f} us/step")

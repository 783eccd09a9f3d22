#!/usr/bin/env python
"""Raw pinned-memory copy bandwidth of this box's PCIe link, the ceiling of bench.py's `e2e` number: one 107 MB device->host
copy (the g + Jacobian slices of 65,536 four-contact instances), one 20 MB host->device copy, and both at once on two streams."""
import torch

torch.cuda.set_device(0)
d2h_bytes, h2d_bytes = 106954752, 20447232
dev_out = torch.empty(d2h_bytes // 8, dtype=torch.float64, device="cuda")
host_out = torch.empty(d2h_bytes // 8, dtype=torch.float64).pin_memory()
dev_in = torch.empty(h2d_bytes // 8, dtype=torch.float64, device="cuda")
host_in = torch.empty(h2d_bytes // 8, dtype=torch.float64).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timed(lambda: host_out.copy_(dev_out, non_blocking=True))
print(f"D2H 107 MB pinned: {ms:.3f} ms = {d2h_bytes / ms / 1e6:.1f} GB/s")
ms = timed(lambda: dev_in.copy_(host_in, non_blocking=True))
print(f"H2D 20 MB pinned: {ms:.3f} ms = {h2d_bytes / ms / 1e6:.1f} GB/s")


def both():
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        host_out.copy_(dev_out, non_blocking=True)
    with torch.cuda.stream(s2):
        dev_in.copy_(host_in, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)


ms = timed(both)
print(f"both at once: {ms:.3f} ms per pair = {d2h_bytes / ms / 1e6:.1f} GB/s D2H + {h2d_bytes / ms / 1e6:.1f} GB/s H2D")

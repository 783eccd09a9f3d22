"""Instance sharding across the GPUs of one box.

Instances are independent, so rank r of G evaluates the contiguous range [b[r], b[r+1]) with the full
parameter set replicated (parameters are kernel arguments).  There is NO collective on the evaluation path.
The only communication is the optional final gather of per-rank output slices for a single consumer,
`gather_outputs`, which uses torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) and is
timed separately from the evaluation by bench.py.
"""
from __future__ import annotations

import numpy as np

GRANULE = 32  # shard boundaries fall on warp-tile boundaries so every rank's slices stay 16-byte aligned


def shard_bounds(num_instances: int, world_size: int, granule: int = GRANULE) -> np.ndarray:
    """Offsets b[0..G]: rank r owns [b[r], b[r+1]).  Balanced to within one granule; covers every instance once."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    blocks = (num_instances + granule - 1) // granule
    b = np.array([min(num_instances, granule * ((blocks * r) // world_size)) for r in range(world_size + 1)], dtype=np.int64)
    b[-1] = num_instances
    return b


def local_range(num_instances: int, rank: int, world_size: int, granule: int = GRANULE):
    b = shard_bounds(num_instances, world_size, granule)
    return int(b[rank]), int(b[rank + 1])


def evaluate_local(problem, x_local, **kw):
    """Evaluate this rank's shard (x_local holds only the local instances).  One kernel launch, no communication."""
    return problem.eval(x_local, **kw)


def gather_outputs(local, num_instances, group=None):
    """All-gather instance-major per-rank outputs {name: (n_local, len) tensor} into full (num_instances, len)
    tensors on every rank.  Shards may differ in size by one granule; they are padded to the largest shard for the
    collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    b = shard_bounds(num_instances, world)
    biggest = int(np.max(np.diff(b)))
    out = {}
    for name, t in local.items():
        if t is None:
            out[name] = None
            continue
        t2 = t.reshape(t.shape[0], -1)
        pad = torch.zeros((biggest, t2.shape[1]), dtype=t2.dtype, device=t2.device)
        pad[: t2.shape[0]] = t2
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        full = torch.cat([parts[r][: int(b[r + 1] - b[r])] for r in range(world)], dim=0)
        out[name] = full.reshape((num_instances,) + tuple(t.shape[1:]))
    return out

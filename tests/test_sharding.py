"""The N > 1 path on CPU: shard arithmetic, and a world_size-2 gloo run in which every rank evaluates its
shard (with the oracle standing in for the GPU) and the gathered result equals the single-process one."""
import os
import socket

import numpy as np
import pytest

from centroidalplanner_b200 import sharding


@pytest.mark.parametrize("N", [0, 1, 31, 32, 33, 1000, 65536, 1 << 20, (1 << 20) + 5])
@pytest.mark.parametrize("G", [1, 2, 3, 4, 8])
def test_shards_cover_every_instance_exactly_once(N, G):
    b = sharding.shard_bounds(N, G)
    assert b[0] == 0 and b[-1] == N and (np.diff(b) >= 0).all()
    assert all(int(v) % sharding.GRANULE == 0 for v in b[:-1])        # aligned starts
    if N >= G * sharding.GRANULE:
        assert np.diff(b).max() - np.diff(b).min() <= 2 * sharding.GRANULE  # balanced
    for r in range(G):
        assert sharding.local_range(N, r, G) == (int(b[r]), int(b[r + 1]))


def _worker(rank, world, port, N, tmp):
    import torch
    import torch.distributed as dist

    from helpers import make_pair

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _, o, gen = make_pair("ground8")
        x = gen(N)                                   # same seeded batch on every rank
        lo, hi = sharding.local_range(N, rank, world)
        ev = o.eval_batch(x[lo:hi], want=("g", "jac"))
        local = {"g": torch.from_numpy(ev["g"]), "jac": torch.from_numpy(ev["jac"]), "cost": None}
        full = sharding.gather_outputs(local, N)
        if rank == 0:
            ref = o.eval_batch(x, want=("g", "jac"))
            ok = np.array_equal(full["g"].numpy(), ref["g"]) and np.array_equal(full["jac"].numpy(), ref["jac"]) and full["cost"] is None
            open(os.path.join(tmp, "ok"), "w").write("1" if ok else "0")
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather(tmp_path):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 1000, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok").read() == "1"

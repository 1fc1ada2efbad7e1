"""N>1 path on CPU: world_size-2 gloo processes shard a batch of instances,
evaluate their blocks and all_gather the values (no GPU needed)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pycollo_b200.parallel import InstanceSharder, shard_range, shard_sizes


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 4096, 4099):
        for w in (1, 2, 4, 8):
            ranges = [shard_range(n, w, r) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges[:-1], ranges[1:]))
            sizes = shard_sizes(n, w)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        X = rng.standard_normal((n, 5))

        def evaluate(x, lam, sigma):          # stand-in for the per-rank CUDA engine
            return {"jac": np.cumsum(x, axis=1) * 2.0}

        sh = InstanceSharder(n, evaluate)
        assert (sh.lo, sh.hi) == shard_range(n, world, rank)
        local = sh.run_local(X)["jac"]
        assert local.shape == (sh.hi - sh.lo, 5)
        full = sh.gather(local)
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)      # the bench's max-over-ranks timing
        if rank == 0:
            q.put((full, float(t)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_instance_sharding():
    n, world = 11, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, 5))
    np.testing.assert_array_equal(full, np.cumsum(X, axis=1) * 2.0)
    assert tmax == 2.0

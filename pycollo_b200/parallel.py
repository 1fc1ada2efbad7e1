"""Multi-GPU partitioning of independent problem instances.

The callback path shards naturally over *independent instances* -- multi-start
or parameter-sweep iterates of one problem (BASELINE config 5) and, for the
throughput benchmark, one independent evaluation stream per GPU.  One process
per GPU (``torchrun``); instances are split into contiguous blocks; there is NO
collective on the data path.  Only when a single consumer wants every instance's
values is one ``all_gather`` issued over NCCL/NVLink (or gloo on CPU tensors in
the tests).

Splitting ONE very large mesh over GPUs (BASELINE config 4) is the other natural
partition -- contiguous tile ranges per rank, an ``all_reduce`` of the few
integral/border partial sums and an ``all_gather`` of value slabs; the tile
tables already support it (tiles are independent, ``tile_desc``), the border
pass split it needs is listed under "next" in DESIGN.md.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_items, world_size, rank):
    """Contiguous, balanced [lo, hi) block of ``n_items`` for ``rank``."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad world_size/rank")
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_sizes(n_items, world_size):
    return [shard_range(n_items, world_size, r)[1] - shard_range(n_items, world_size, r)[0]
            for r in range(world_size)]


class InstanceSharder:
    """Evaluate a batch of independent iterates split across ranks.

    ``evaluate(x_block, lam_block, sigma_block) -> dict of 2-D arrays`` is the
    per-rank evaluator (the CUDA engine with ``batch = block size`` on a GPU
    rank; anything with the same signature in CPU tests).
    """

    def __init__(self, n_instances, evaluate, world_size=None, rank=None):
        import torch.distributed as dist
        self.dist = dist
        if world_size is None:
            world_size = dist.get_world_size() if dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
        self.world_size, self.rank = int(world_size), int(rank)
        self.n = int(n_instances)
        self.lo, self.hi = shard_range(self.n, self.world_size, self.rank)
        self.evaluate = evaluate

    def local_slice(self, arr):
        return arr[self.lo:self.hi]

    def run_local(self, X, LAM=None, SIGMA=None):
        """Evaluate this rank's block of the (global) input arrays."""
        x = self.local_slice(X)
        lam = None if LAM is None else self.local_slice(LAM)
        sig = None if SIGMA is None else self.local_slice(SIGMA)
        return self.evaluate(x, lam, sig)

    def gather(self, local, device=None):
        """all_gather a per-rank (block, m) array into the (n, m) global one."""
        import torch
        if self.world_size == 1 or not self.dist.is_initialized():
            return np.asarray(local)
        local_t = local if isinstance(local, torch.Tensor) else \
            torch.from_numpy(np.ascontiguousarray(local))
        if device is not None:
            local_t = local_t.to(device)
        m = local_t.shape[1]
        sizes = shard_sizes(self.n, self.world_size)
        pad = max(sizes)
        buf = torch.zeros((pad, m), dtype=local_t.dtype, device=local_t.device)
        buf[:local_t.shape[0]] = local_t
        out = [torch.empty_like(buf) for _ in range(self.world_size)]
        self.dist.all_gather(out, buf)
        full = torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)
        return full.cpu().numpy()

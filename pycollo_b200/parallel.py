"""Multi-GPU partitioning of independent problem instances.

The callback path shards naturally over *independent instances* -- multi-start
or parameter-sweep iterates of one problem (BASELINE config 5) and, for the
throughput benchmark, one independent evaluation stream per GPU.  One process
per GPU (``torchrun``); instances are split into contiguous blocks; there is NO
collective on the data path.  Only when a single consumer wants every instance's
values is one ``all_gather`` issued over NCCL/NVLink (or gloo on CPU tensors in
the tests).

Splitting ONE very large mesh over GPUs (BASELINE config 4) is the other natural
partition (``MeshSharder``): every rank builds the same engine, restricts it to
a contiguous range of tiles (``pcx_set_shard``), evaluates its slabs of the
value arrays, and the only exchange is one NCCL ``all_reduce`` of a few dozen
doubles -- the quadrature partial sums and end-node values the border slots
need -- between the two stages of the border pass.  x and lam are replicated
(each rank reads them from its own pinned host buffer or a broadcast).
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


class _DeviceArray:
    """CUDA-array-interface view of a raw device pointer (for torch.as_tensor)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = dict(shape=(int(n),), typestr="<f8",
                                             data=(int(ptr), False), version=2)


class MeshSharder:
    """One mesh split over the ranks by contiguous tile ranges (config 4).

    ``evaluate`` = stage 1 (this rank's tiles; its slab of every selected value
    array, slots disjoint between ranks) -> ``all_reduce(sum)`` of the border
    exchange buffer over NCCL -> stage 2 (border slots).  With ``border_rank``
    set, only that rank writes the O(1) border slots, so that a sum over ranks of
    zero-initialised arrays reassembles the full vectors exactly (each slot has
    exactly one writer); with ``border_rank=None`` every rank writes them.
    """

    def __init__(self, engine, world_size=None, rank=None, border_rank=None, fused=False):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        if world_size is None:
            world_size = dist.get_world_size() if dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
        self.world_size, self.rank = int(world_size), int(rank)
        self.engine = engine
        self.border_rank = border_rank
        self.fused = bool(fused) and self.world_size > 1
        self.lo, self.hi = shard_range(engine.S.num_tiles, self.world_size, self.rank)
        engine.set_shard(self.lo, self.hi)
        ptr, n = engine.shard_buffer()
        self.xbuf = torch.as_tensor(_DeviceArray(ptr, n), device=f"cuda:{engine.device}")
        if self.fused:
            # exchange fused into the kernel over peer memory (NVLink): the border
            # rank owns the buffer, the others map it through CUDA IPC
            br = 0 if border_rank is None else int(border_rank)
            self.border_rank = br
            box = [engine.exchange_alloc(self.world_size) if self.rank == br else None]
            dist.broadcast_object_list(box, src=br)
            engine.exchange_attach(self.rank, self.world_size, br,
                                   None if self.rank == br else box[0])
            dist.barrier()

    def evaluate(self, what, x, lam=None, sigma=None, f=None, grad=None, c=None, dy=None,
                 jac=None, hess=None):
        """Device tensors in, device tensors (full-size, this rank's slab filled) out;
        everything is enqueued on torch's current stream."""
        st = self.torch.cuda.current_stream().cuda_stream
        self.engine.eval_ptr(what, x, lam=lam, sigma=sigma, f=f, grad=grad, c=c, dy=dy,
                             jac=jac, hess=hess, stream=st)
        if self.world_size == 1 or self.fused:
            return     # unsharded: pcx_eval did the border pass; fused: the kernel exchanged
        self.dist.all_reduce(self.xbuf)
        if self.border_rank is None or self.border_rank == self.rank:
            self.engine.apply_border(what, x, lam=lam, sigma=sigma, f=f, grad=grad, c=c,
                                     jac=jac, hess=hess, stream=st)

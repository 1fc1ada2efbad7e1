"""Two (or more) processes, one GPU each: the fused peer-memory border exchange
across processes (CUDA IPC mapping of the border rank's buffer, system-scope
release/acquire flags, double-buffered shares with back-pressure).

Run by tests/test_sharding.py::test_gpu_fused_exchange_across_processes as
    python -m torch.distributed.run --nproc-per-node 2 tests/dist_fused_exchange.py
Many evaluations with DIFFERENT iterates are enqueued back to back with no host
synchronisation in between (a fast rank runs ahead of the border rank); every
one of them must reproduce the unsharded evaluation exactly."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

from examples import problems
from examples.cases import lower_case
from pycollo_b200 import engine as E
from pycollo_b200.parallel import MeshSharder


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    low, _, scal = lower_case(problems.multiphase_sliding_mass(), "lobatto", 40, 4, seed=3,
                              max_tile_nodes=16)
    S = low.S
    what = E.EVAL_C | E.EVAL_JAC | E.EVAL_HESS | E.EVAL_F | E.EVAL_GRAD
    eng = E.Engine(S, low.layouts, low.header, device=local)
    eng.set_scaling(*scal)
    n_evals = 40
    rng = np.random.default_rng(11)                     # same iterates on every rank
    xs = [torch.from_numpy(rng.uniform(-0.5, 0.5, S.num_x)).to(dev) for _ in range(n_evals)]
    lams = [torch.from_numpy(rng.standard_normal(S.num_c)).to(dev) for _ in range(n_evals)]
    sig = torch.tensor([0.7], dtype=torch.float64, device=dev)
    z = lambda n: torch.zeros(n, dtype=torch.float64, device=dev)
    mk = lambda: dict(f=z(1), grad=z(S.num_x), c=z(S.num_c), jac=z(S.nnz_g), hess=z(S.nnz_h))
    ref = [mk() for _ in range(n_evals)]
    st = torch.cuda.current_stream().cuda_stream
    for x, lam, o in zip(xs, lams, ref):                # unsharded, on every rank
        eng.eval_ptr(what, x, lam=lam, sigma=sig, stream=st, **o)
    torch.cuda.synchronize()
    sh = MeshSharder(eng, world, rank, border_rank=0, fused=True)
    out = [mk() for _ in range(n_evals)]
    dist.barrier()
    if rank != 0:                                       # let the non-border ranks run ahead
        pass
    else:
        torch.cuda._sleep(200_000_000)                  # ~0.1 s: the border rank starts late
    for x, lam, o in zip(xs, lams, out):                # no synchronisation in between
        sh.evaluate(what, x, lam=lam, sigma=sig, **o)
    torch.cuda.synchronize()
    assert eng.status() == 0, "exchange timed out"
    worst, bad = 0.0, []
    for i, (o, r) in enumerate(zip(out, ref)):
        for k in ("c", "jac", "hess", "grad", "f"):
            t = o[k].clone()
            dist.all_reduce(t)                          # sum of disjoint slabs (+ border rank's slots)
            d = float((t - r[k]).abs().max())
            s = float(r[k].abs().max()) or 1.0
            worst = max(worst, d / s)
            if d / s > 1e-13:
                bad.append((i, k, d / s, int((t - r[k]).abs().argmax())))
    if rank == 0 and bad:
        print("deviating (evaluation, output, rel. deviation, slot):", bad[:24], flush=True)
    assert worst <= 1e-13, worst
    if rank == 0:
        print(f"fused exchange across {world} processes: {n_evals} evaluations, "
              f"worst deviation from the unsharded result {worst:.1e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

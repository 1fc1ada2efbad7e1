"""Config 5: batched multi-start sweep -- B independent instances of a small
problem (10 sections x 4 nodes = 31 nodes) evaluated in ONE launch (grid.y = B).
Prints instance-evaluations/s for fused G+H and the algorithmic GB/s."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import build_case
from pycollo_b200 import engine as E
from examples import problems as examples


RANK = int(os.environ.get("RANK", 0)); WORLD = int(os.environ.get("WORLD_SIZE", 1))
LOCAL = int(os.environ.get("LOCAL_RANK", 0))


def run(problem, batch, threads, steps=30):
    """`batch` instances in total; under torchrun every rank takes its contiguous
    block (parallel.shard_range), no collective on the data path."""
    import torch.distributed as dist
    from pycollo_b200.parallel import shard_range
    total = batch
    lo, hi = shard_range(total, WORLD, RANK)
    batch = hi - lo
    torch.cuda.set_device(LOCAL)
    low, _, scal = build_case(getattr(examples, problem)(), "lobatto", 10, 4, seed=0,
                              oracle=False, threads=threads, max_tile_nodes=threads)
    S = low.S
    eng = E.Engine(S, low.layouts, low.header, batch=batch, device=LOCAL)
    eng.set_scaling(*scal)
    per = 8 * (S.num_x + S.nnz_g) + 8 * (S.num_x + S.num_c + S.nnz_h)
    R = max(2, int(np.ceil(300e6 / (per * batch))))
    g = torch.Generator(device="cuda").manual_seed(0)
    xs = [torch.rand(batch, S.num_x, dtype=torch.float64, device="cuda", generator=g) - 0.5 for _ in range(R)]
    ls = [torch.randn(batch, S.num_c, dtype=torch.float64, device="cuda", generator=g) for _ in range(R)]
    js = [torch.empty(batch, S.nnz_g, dtype=torch.float64, device="cuda") for _ in range(R)]
    hs = [torch.empty(batch, S.nnz_h, dtype=torch.float64, device="cuda") for _ in range(R)]
    st = torch.cuda.current_stream().cuda_stream
    what = E.EVAL_JAC | E.EVAL_HESS
    for i in range(4):
        eng.eval_ptr(what, xs[i % R], lam=ls[i % R], jac=js[i % R], hess=hs[i % R], stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        eng.eval_ptr(what, xs[i % R], lam=ls[i % R], jac=js[i % R], hess=hs[i % R], stream=st)
    e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / steps
    if WORLD > 1:
        t = torch.tensor([us], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)           # time = max over ranks
        us = float(t)
    batch = total
    if RANK == 0:
      print(json.dumps(dict(problem=problem, n_gpus=WORLD, batch=batch, threads=threads, tiles_per_instance=S.num_tiles,
                          num_x=S.num_x, nnz_G=S.nnz_g, nnz_H=S.nnz_h, us_per_launch=round(us, 2),
                          instance_evals_per_s=round(batch / us * 1e6), gbs=round(per * batch / us / 1e3, 1))), flush=True)


if __name__ == "__main__":
    if WORLD > 1:
        import torch.distributed as dist
        torch.cuda.set_device(LOCAL)
        dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
    for prob in ("cart_pole_swing_up", "hypersensitive"):
        for thr in (32, 128):
            try:
                run(prob, 4096, thr)
            except Exception as exc:
                print("FAILED", prob, thr, repr(exc)[:300], flush=True)
    if WORLD > 1:
        import torch.distributed as dist
        dist.destroy_process_group()

"""What ONE rank of an N-way sharded Delta III mesh costs, measured on one GPU: the
engine is restricted to rank r's tile range (stage 1 only: no exchange) and timed, for
several tilings.  python tools/shard_sim.py [world] [K]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from examples import problems
from examples.cases import lower_case
from pycollo_b200 import engine as E
from pycollo_b200.parallel import shard_range

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
K = int(sys.argv[2]) if len(sys.argv) > 2 else 83333
dev = torch.device("cuda")
what = E.EVAL_JAC | E.EVAL_HESS
for label, kw in (("1gpu-tiling", {}), ("sm148xW", dict(sm_count=148 * world)),
                  ("W x 4/SM", dict(sm_count=148 * world, tiles_per_sm=4)),
                  ("W x 6/SM", dict(sm_count=148 * world, tiles_per_sm=6)),
                  ("W x 8/SM", dict(sm_count=148 * world, tiles_per_sm=8)),
                  ("W x 9/SM", dict(sm_count=148 * world, tiles_per_sm=9)),
                  ("W x 10/SM", dict(sm_count=148 * world, tiles_per_sm=10))):
    low, _, scal = lower_case(problems.delta_iii_launch_vehicle(), "lobatto", K, 4, seed=0, **kw)
    S = low.S
    eng = E.Engine(S, low.layouts, low.header, structure=False)
    eng.set_scaling(*scal)
    g = torch.Generator(device=dev).manual_seed(0)
    x = 0.1 + 0.3 * torch.rand(S.num_x, dtype=torch.float64, device=dev, generator=g)
    lam = torch.randn(S.num_c, dtype=torch.float64, device=dev, generator=g)
    sets = [dict(x=x, lam=lam, jac=torch.empty(S.nnz_g, dtype=torch.float64, device=dev),
                 hess=torch.empty(S.nnz_h, dtype=torch.float64, device=dev))]
    args = eng.make_args(sets)
    st = torch.cuda.current_stream().cuda_stream
    info = eng.variant_info(what)
    eng.eval_many(what, args, 3, stream=st, gate=False, timed=False)
    torch.cuda.synchronize()
    full = eng.eval_many(what, args, 10, stream=st, gate=True, timed=True) / 10
    res = {}
    for r in (0, world // 2, world - 1):
        lo, hi = shard_range(S.num_tiles, world, r)
        eng.set_shard(lo, hi)
        eng.eval_many(what, args, 3, stream=st, gate=False, timed=False)
        torch.cuda.synchronize()
        res[r] = round(1e3 * eng.eval_many(what, args, 20, stream=st, gate=True, timed=True) / 20, 1)
    print(json.dumps(dict(label=label, tiles=int(S.num_tiles), per_rank=int(S.num_tiles // world),
                          max_tile_nodes=int(S.max_tile_nodes), info=info,
                          full_ms=round(full, 4), rank_us=res)), flush=True)
    del eng, sets, args
    torch.cuda.empty_cache()

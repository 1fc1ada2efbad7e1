"""Per-range times (us, fused G+H) of a W-way sharded Delta III mesh on ONE GPU for given tiles per SM:
   python tools/tiling_ranges.py W:m [W:m ...]     (m = tiles per SM per rank, no rounding)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from examples import problems
from examples.cases import lower_case
from pycollo_b200 import engine as E
from pycollo_b200.parallel import shard_range

K = 83333
dev = torch.device("cuda")
what = E.EVAL_JAC | E.EVAL_HESS
st = torch.cuda.current_stream().cuda_stream
for spec in sys.argv[1:]:
    W, m = (int(v) for v in spec.split(":"))
    low, _, scal = lower_case(problems.delta_iii_launch_vehicle(), "lobatto", K, 4, seed=0,
                              sm_count=148 * W, tiles_per_sm=m)
    S = low.S
    eng = E.Engine(S, low.layouts, low.header, structure=False)
    eng.set_scaling(*scal)
    g = torch.Generator(device=dev).manual_seed(0)
    x = 0.1 + 0.3 * torch.rand(S.num_x, dtype=torch.float64, device=dev, generator=g)
    lam = torch.randn(S.num_c, dtype=torch.float64, device=dev, generator=g)
    jac = torch.zeros(S.nnz_g, dtype=torch.float64, device=dev)
    hes = torch.zeros(S.nnz_h, dtype=torch.float64, device=dev)
    args = eng.make_args([dict(x=x, lam=lam, jac=jac, hess=hes)])
    row = []
    for r in range(W):
        eng.set_shard(*shard_range(S.num_tiles, W, r))
        eng.eval_many(what, args, 3, stream=st, gate=False, timed=False)
        torch.cuda.synchronize()
        row.append(round(1e3 * eng.eval_many(what, args, 12, stream=st, gate=True, timed=True) / 12, 1))
    print(json.dumps(dict(W=W, tiles_per_sm=m, tiles=int(S.num_tiles), max_tile_nodes=int(S.max_tile_nodes),
                          jac_hess_us=row)), flush=True)
    del eng, args, jac, hes, x, lam
    torch.cuda.empty_cache()

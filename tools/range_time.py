"""us per tile range of a W-way sharded Delta III mesh, every range timed on ONE GPU.
  python tools/range_time.py W K [K ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from examples import problems
from examples.cases import lower_case
from pycollo_b200 import engine as E
from pycollo_b200.parallel import shard_range

W = int(sys.argv[1])
dev = torch.device("cuda")
what = E.EVAL_JAC | E.EVAL_HESS
for K in [int(a) for a in sys.argv[2:]]:
    low, _, scal = lower_case(problems.delta_iii_launch_vehicle(), "lobatto", K, 4, seed=0, sm_count=148 * W)
    S = low.S
    eng = E.Engine(S, low.layouts, low.header, structure=False)
    eng.set_scaling(*scal)
    g = torch.Generator(device=dev).manual_seed(0)
    x = 0.1 + 0.3 * torch.rand(S.num_x, dtype=torch.float64, device=dev, generator=g)
    lam = torch.randn(S.num_c, dtype=torch.float64, device=dev, generator=g)
    pad = int(os.environ.get("PAD", 0))
    jac = torch.zeros(S.nnz_g + pad, dtype=torch.float64, device=dev)[pad:]
    hes = torch.zeros(S.nnz_h + pad, dtype=torch.float64, device=dev)[pad:]
    args = eng.make_args([dict(x=x, lam=lam, jac=jac, hess=hes)])
    st = torch.cuda.current_stream().cuda_stream
    res = []
    for mode in (E.EVAL_JAC | E.EVAL_HESS, E.EVAL_JAC, E.EVAL_HESS):
        row = []
        for r in range(W):
            eng.set_shard(*shard_range(S.num_tiles, W, r))
            eng.eval_many(mode, args, 3, stream=st, gate=False, timed=False)
            torch.cuda.synchronize()
            row.append(round(1e3 * eng.eval_many(mode, args, 20, stream=st, gate=True, timed=True) / 20, 1))
        res.append(row)
    print(json.dumps(dict(K=K, W=W, tiles=int(S.num_tiles), jac_hess_us=res[0], jac_us=res[1], hess_us=res[2])), flush=True)
    del eng, args, jac, hes
    torch.cuda.empty_cache()

"""Delta III, 10^6 nodes: the tiling with the corrected residency estimate (3 CTAs per SM for
the two-pass body: 54 tiles per SM = 18 whole waves; 9 per SM on an 8-way sharded mesh) against
the previous one (56 / 8 per SM, built for 4 CTAs per SM), values compared bit for bit, and the
per-range times of the 8-way sharded mesh on one GPU.   python tools/tiling_ab.py [K]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from examples import problems
from examples.cases import lower_case
from pycollo_b200 import engine as E
from pycollo_b200.parallel import shard_range

K = int(sys.argv[1]) if len(sys.argv) > 1 else 83333
dev = torch.device("cuda")
what = E.EVAL_JAC | E.EVAL_HESS
st = torch.cuda.current_stream().cuda_stream
ref = None
for label, kw in (("new", {}), ("old", dict(tiles_per_sm=56))):
    low, _, scal = lower_case(problems.delta_iii_launch_vehicle(), "lobatto", K, 4, seed=0, **kw)
    S = low.S
    eng = E.Engine(S, low.layouts, low.header, structure=False)
    eng.set_scaling(*scal)
    g = torch.Generator(device=dev).manual_seed(0)
    x = 0.1 + 0.3 * torch.rand(S.num_x, dtype=torch.float64, device=dev, generator=g)
    lam = torch.randn(S.num_c, dtype=torch.float64, device=dev, generator=g)
    sets = [dict(x=x, lam=lam, jac=torch.zeros(S.nnz_g, dtype=torch.float64, device=dev),
                 hess=torch.zeros(S.nnz_h, dtype=torch.float64, device=dev)) for _ in range(2)]
    args = eng.make_args(sets)
    eng.eval_many(what, args, 4, stream=st, gate=False, timed=False)
    torch.cuda.synchronize()
    ms = eng.eval_many(what, args, 10, stream=st, gate=True, timed=True) / 10
    alg = 8 * (S.num_x + S.nnz_g) + 8 * (S.num_x + S.num_c + S.nnz_h)
    out = dict(tiling=label, tiles=int(S.num_tiles), max_tile_nodes=int(S.max_tile_nodes),
               ms_per_eval=round(ms, 4), frac=round(alg / (ms * 1e-3) / 6553e9, 4),
               blocks_per_sm=eng.variant_info(what)["blocks_per_sm"])
    if ref is None:
        ref = (sets[0]["jac"].clone(), sets[0]["hess"].clone())
    else:
        out["jac_bitwise_equal"] = bool(torch.equal(ref[0], sets[0]["jac"]))
        out["hess_bitwise_equal"] = bool(torch.equal(ref[1], sets[0]["hess"]))
        out["max_abs_diff"] = [float((ref[0] - sets[0]["jac"]).abs().max()),
                               float((ref[1] - sets[0]["hess"]).abs().max())]
    print(json.dumps(out), flush=True)
    del eng, sets, args
    torch.cuda.empty_cache()
W = 8
low, _, scal = lower_case(problems.delta_iii_launch_vehicle(), "lobatto", K, 4, seed=0, sm_count=148 * W)
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
    row.append(round(1e3 * eng.eval_many(what, args, 20, stream=st, gate=True, timed=True) / 20, 1))
# the union of the ranges' slabs against the unsharded evaluation: everything but the border
# entries (endpoint rows / block: applied by pcx_apply_border or the fused exchange, not here)
print(json.dumps(dict(W=W, tiles=int(S.num_tiles), max_tile_nodes=int(S.max_tile_nodes), jac_hess_us=row,
                      slab_entries_differing_from_unsharded=[int((jac != ref[0]).sum()), int((hes != ref[1]).sum())],
                      border_entries=int(len(S.border_grp)))), flush=True)

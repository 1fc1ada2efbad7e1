"""One launch per tile range of a W-way sharded Delta III mesh on ONE GPU (for ncu:
do the ranges execute the same number of instructions / move the same bytes?).
  python tools/range_probe.py [W] [K]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from examples import problems
from examples.cases import lower_case
from pycollo_b200 import engine as E
from pycollo_b200.parallel import shard_range

W = int(sys.argv[1]) if len(sys.argv) > 1 else 4
K = int(sys.argv[2]) if len(sys.argv) > 2 else 83333
dev = torch.device("cuda")
what = E.EVAL_JAC | E.EVAL_HESS
low, _, scal = lower_case(problems.delta_iii_launch_vehicle(), "lobatto", K, 4, seed=0, sm_count=148 * W)
S = low.S
eng = E.Engine(S, low.layouts, low.header, structure=False)
eng.set_scaling(*scal)
g = torch.Generator(device=dev).manual_seed(0)
x = 0.1 + 0.3 * torch.rand(S.num_x, dtype=torch.float64, device=dev, generator=g)
lam = torch.randn(S.num_c, dtype=torch.float64, device=dev, generator=g)
jac = torch.zeros(S.nnz_g, dtype=torch.float64, device=dev)
hes = torch.zeros(S.nnz_h, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
for r in range(W):
    lo, hi = shard_range(S.num_tiles, W, r)
    eng.set_shard(lo, hi)
    print("range", r, lo, hi, "phases", sorted(set(S.tile_phase[lo:hi].tolist())),
          "nodes", int(S.tile_nodes[lo:hi].sum()), flush=True)
    for _ in range(2):
        eng.eval_ptr(what, x, lam=lam, jac=jac, hess=hes, stream=st)
    torch.cuda.synchronize()

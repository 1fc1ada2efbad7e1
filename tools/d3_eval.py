"""BASELINE config 4 on ONE GPU: Delta III, 4 phases x K sections x 4 nodes (default
~10^6 nodes), fused G+H, a few launches (for ncu / timing).  python tools/d3_eval.py [K] [steps]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from examples import problems
from examples.cases import lower_case
from pycollo_b200 import engine as E

K = int(sys.argv[1]) if len(sys.argv) > 1 else 83333
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
PROBLEM = sys.argv[3] if len(sys.argv) > 3 else "delta_iii_launch_vehicle"      # any examples.problems name
low, _, scal = lower_case(getattr(problems, PROBLEM)(), "lobatto", K, 4, seed=0)
S = low.S
eng = E.Engine(S, low.layouts, low.header, structure=False)
eng.set_scaling(*scal)
what = {"jh": E.EVAL_JAC | E.EVAL_HESS, "j": E.EVAL_JAC, "h": E.EVAL_HESS}[os.environ.get("WHAT", "jh")]
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
x = 0.1 + 0.3 * torch.rand(S.num_x, dtype=torch.float64, device=dev, generator=g)
lam = torch.randn(S.num_c, dtype=torch.float64, device=dev, generator=g)
sets = [dict(x=x, lam=lam, jac=torch.empty(S.nnz_g, dtype=torch.float64, device=dev),
             hess=torch.empty(S.nnz_h, dtype=torch.float64, device=dev)) for _ in range(2)]
args = eng.make_args(sets)
st = torch.cuda.current_stream().cuda_stream
eng.eval_many(what, args, 3, stream=st, gate=False, timed=False)
torch.cuda.synchronize()
ms = eng.eval_many(what, args, steps, stream=st, gate=True, timed=True) / steps
alg = (8 * (S.num_x + S.nnz_g) if what & E.EVAL_JAC else 0) + (8 * (S.num_x + S.num_c + S.nnz_h) if what & E.EVAL_HESS else 0)
print(json.dumps(dict(workload="delta_iii 4 phases" if PROBLEM.startswith("delta") else PROBLEM, nodes=int(sum(t.N for t in S.ph)), tiles=int(S.num_tiles),
                      threads=int(S.threads), ms_per_eval=round(ms, 4), algorithmic_GBs=round(alg / ms / 1e6, 1),
                      frac=round(alg / (ms * 1e-3) / 6553e9, 4),
                      what=os.environ.get("WHAT", "jh"),
                      env={k: v for k, v in os.environ.items() if k.startswith("PCX_")})), flush=True)

"""Fused G+H launch vs separate Jacobian and Hessian launches (big expression bodies)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import build_case
from pycollo_b200 import engine as E
from examples import problems as examples

problem = sys.argv[1] if len(sys.argv) > 1 else "delta_iii_launch_vehicle"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 83333
low, _, scal = build_case(getattr(examples, problem)(), "lobatto", K, 4, seed=0, oracle=False)
S = low.S
eng = E.Engine(S, low.layouts, low.header)
eng.set_scaling(*scal)
g = torch.Generator(device="cuda").manual_seed(0)
x = 0.1 + 0.3 * torch.rand(S.num_x, dtype=torch.float64, device="cuda", generator=g)
lam = torch.randn(S.num_c, dtype=torch.float64, device="cuda", generator=g)
R = 2
jac = [torch.empty(S.nnz_g, dtype=torch.float64, device="cuda") for _ in range(R)]
hes = [torch.empty(S.nnz_h, dtype=torch.float64, device="cuda") for _ in range(R)]
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, steps=20):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / steps


fused = timeit(lambda i: eng.eval_ptr(E.EVAL_JAC | E.EVAL_HESS, x, lam=lam, jac=jac[i % R], hess=hes[i % R], stream=st))
jo = timeit(lambda i: eng.eval_ptr(E.EVAL_JAC, x, jac=jac[i % R], stream=st))
ho = timeit(lambda i: eng.eval_ptr(E.EVAL_HESS, x, lam=lam, hess=hes[i % R], stream=st))
both = timeit(lambda i: (eng.eval_ptr(E.EVAL_JAC, x, jac=jac[i % R], stream=st),
                         eng.eval_ptr(E.EVAL_HESS, x, lam=lam, hess=hes[i % R], stream=st)))
print(json.dumps(dict(problem=problem, nodes=int(sum(t.N for t in S.ph)), fused_us=round(fused, 1),
                      jac_us=round(jo, 1), hess_us=round(ho, 1), jac_then_hess_us=round(both, 1))))

"""Fixed-cost amortisation of the fused G+H kernel: batch of independent iterates
per launch (config 5 style) and a 10x larger mesh (config 4 style), cart-pole.
Prints us per evaluation and algorithmic GB/s (8*(num_x+nnz_G) + 8*(num_x+num_c+nnz_H))."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import build_case
from pycollo_b200 import engine as E
from examples import problems as examples


def run(sections, batch, steps=60, min_blocks=None, tps=None):
    low, _, scal = build_case(examples.cart_pole_swing_up(), "lobatto", sections, 4, seed=0,
                              unit_scaling=True, oracle=False, tiles_per_sm=tps)
    S = low.S
    eng = E.Engine(S, low.layouts, low.header, batch=batch, min_blocks=min_blocks)
    eng.set_scaling(*scal)
    per_eval = 8 * (S.num_x + S.nnz_g) + 8 * (S.num_x + S.num_c + S.nnz_h)
    R = max(2, int(np.ceil(300e6 / (per_eval * batch))))       # ring > L2
    g = torch.Generator(device="cuda").manual_seed(0)
    xs = [torch.rand(batch, S.num_x, dtype=torch.float64, device="cuda", generator=g) - 0.5 for _ in range(R)]
    ls = [torch.randn(batch, S.num_c, dtype=torch.float64, device="cuda", generator=g) for _ in range(R)]
    js = [torch.empty(batch, S.nnz_g, dtype=torch.float64, device="cuda") for _ in range(R)]
    hs = [torch.empty(batch, S.nnz_h, dtype=torch.float64, device="cuda") for _ in range(R)]
    st = torch.cuda.current_stream().cuda_stream
    what = E.EVAL_JAC | E.EVAL_HESS
    for i in range(6):
        eng.eval_ptr(what, xs[i % R], lam=ls[i % R], jac=js[i % R], hess=hs[i % R], stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        eng.eval_ptr(what, xs[i % R], lam=ls[i % R], jac=js[i % R], hess=hs[i % R], stream=st)
    e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / steps
    print(json.dumps(dict(nodes=3 * sections + 1, batch=batch, tiles=S.num_tiles, min_blocks=min_blocks,
                          us_per_launch=round(us, 2), us_per_eval=round(us / batch, 2),
                          gbs=round(per_eval * batch / us / 1e3, 1),
                          frac=round(per_eval * batch / us / 1e3 / 6553.0, 3))), flush=True)
    del eng


if __name__ == "__main__":
    cfgs = json.loads(sys.argv[1]) if len(sys.argv) > 1 else \
        [(33333, 1), (33333, 2), (33333, 4), (33333, 8), (333333, 1), (333333, 2)]
    for c in cfgs:
        try:
            run(*c)
        except Exception as exc:
            print("FAILED", c, repr(exc)[:300], flush=True)

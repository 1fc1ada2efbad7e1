"""Config 1: brachistochrone, default mesh (10 sections x 4 nodes, num_x=125, num_c=90):
latency of each callback through the C ABI (device-resident and host arrays)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import build_case
from pycollo_b200 import engine as E
from examples import problems as examples

for method in ("lobatto", "radau"):
    low, _, scal = build_case(examples.brachistochrone(), method, 10, 4, seed=0, oracle=False)
    S = low.S
    eng = E.Engine(S, low.layouts, low.header)
    eng.set_scaling(*scal)
    rng = np.random.default_rng(0)
    xh, lh = rng.uniform(-0.5, 0.5, S.num_x), rng.standard_normal(S.num_c)
    x, lam = torch.from_numpy(xh).cuda(), torch.from_numpy(lh).cuda()
    z = lambda n: torch.empty(n, dtype=torch.float64, device="cuda")
    out = dict(f=z(1), grad=z(S.num_x), c=z(S.num_c), jac=z(S.nnz_g), hess=z(S.nnz_h))
    st = torch.cuda.current_stream().cuda_stream
    res = dict(problem="brachistochrone", quadrature=method, num_x=S.num_x, num_c=S.num_c,
               nnz_G=S.nnz_g, nnz_H=S.nnz_h, threads=S.threads, tiles=S.num_tiles)
    ALL = E.EVAL_F | E.EVAL_GRAD | E.EVAL_C | E.EVAL_JAC | E.EVAL_HESS
    for name, what, kw in (("jac+hess", E.EVAL_JAC | E.EVAL_HESS, dict(jac=out["jac"], hess=out["hess"])),
                           ("all five", ALL, out)):
        for _ in range(20):
            eng.eval_ptr(what, x, lam=lam, **kw, stream=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(500):
            eng.eval_ptr(what, x, lam=lam, **kw, stream=st)
        e1.record(); torch.cuda.synchronize()
        res[f"{name} device us"] = round(1e3 * e0.elapsed_time(e1) / 500, 2)
    t0 = time.perf_counter()
    for _ in range(200):
        eng.eval_host(ALL, xh, lh, 1.0)
    res["all five host-array us"] = round(1e6 * (time.perf_counter() - t0) / 200, 1)
    print(json.dumps(res), flush=True)

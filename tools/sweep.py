"""Tuning sweep of the fused G+H kernel on config 2 (threads, tiles/SM, reg cap)."""
import itertools, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import build_case
from pycollo_b200 import engine as E
from examples import problems as examples

def run(threads, tps, maxreg, extra=None, steps=100):
    os.environ['PCX_NVRTC_EXTRA'] = extra or ''
    low, _, scal = build_case(examples.cart_pole_swing_up(), "lobatto", 33333, 4, seed=0,
                              unit_scaling=True, oracle=False, threads=threads,
                              max_tile_nodes=threads, tiles_per_sm=tps)
    S = low.S
    eng = E.Engine(S, low.layouts, low.header, min_blocks=maxreg)
    eng.set_scaling(*scal)
    R = 6
    rng = np.random.default_rng(0)
    xs = [torch.from_numpy(rng.uniform(-.5, .5, S.num_x)).cuda() for _ in range(R)]
    ls = [torch.from_numpy(rng.standard_normal(S.num_c)).cuda() for _ in range(R)]
    js = [torch.empty(S.nnz_g, dtype=torch.float64, device="cuda") for _ in range(R)]
    hs = [torch.empty(S.nnz_h, dtype=torch.float64, device="cuda") for _ in range(R)]
    st = torch.cuda.current_stream().cuda_stream
    what = E.EVAL_JAC | E.EVAL_HESS
    for i in range(10):
        eng.eval_ptr(what, xs[i % R], lam=ls[i % R], jac=js[i % R], hess=hs[i % R], stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        eng.eval_ptr(what, xs[i % R], lam=ls[i % R], jac=js[i % R], hess=hs[i % R], stream=st)
    e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / steps
    # host-side cost of one call (no sync)
    t0 = time.perf_counter()
    for i in range(steps):
        eng.eval_ptr(what, xs[i % R], lam=ls[i % R], jac=js[i % R], hess=hs[i % R], stream=st)
    host_us = 1e6 * (time.perf_counter() - t0) / steps
    torch.cuda.synchronize()
    # the same launches captured in a CUDA graph (no host launch cost)
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    n_in_graph = 2 * R
    with torch.cuda.stream(side):
        g.capture_begin()
        for i in range(n_in_graph):
            eng.eval_ptr(what, xs[i % R], lam=ls[i % R], jac=js[i % R], hess=hs[i % R],
                         stream=side.cuda_stream)
        g.capture_end()
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    reps = 20
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    graph_us = 1e3 * e0.elapsed_time(e1) / (reps * n_in_graph)
    print(json.dumps(dict(threads=threads, tiles_per_sm=tps, min_blocks=maxreg, tiles=S.num_tiles,
                          extra=extra, nodes=S.max_tile_nodes, us=round(us, 2), host_us=round(host_us, 2),
                          graph_us=round(graph_us, 2),
                          gbs=round(46399696 / graph_us / 1e3, 1))), flush=True)

if __name__ == "__main__":
    cfgs = json.loads(sys.argv[1]) if len(sys.argv) > 1 else \
        [(128, None, None), (128, None, 96), (128, 8, None), (64, None, None), (256, None, None)]
    for c in cfgs:
        try:
            run(*c)
        except Exception as exc:
            print("FAILED", c, repr(exc)[:300], flush=True)

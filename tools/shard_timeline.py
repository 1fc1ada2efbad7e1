"""Per-tile timeline (%globaltimer stamps, debug build) of single tile ranges of the
W-way sharded Delta III mesh on one GPU: why are some ranges slower than others?
  python tools/shard_timeline.py W r [r ...]"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PCX_NVRTC_EXTRA"] = "-DPCX_DEBUG_TIMELINE"
import numpy as np, torch
from examples import problems
from examples.cases import lower_case
from pycollo_b200 import engine as E
from pycollo_b200.parallel import shard_range

W = int(sys.argv[1]); ranges = [int(a) for a in sys.argv[2:]]
dev = torch.device("cuda")
what = E.EVAL_JAC | E.EVAL_HESS
low, _, scal = lower_case(problems.delta_iii_launch_vehicle(), "lobatto", 83333, 4, seed=0, sm_count=148 * W)
S = low.S
eng = E.Engine(S, low.layouts, low.header, structure=False)
eng.set_scaling(*scal)
g = torch.Generator(device=dev).manual_seed(0)
x = 0.1 + 0.3 * torch.rand(S.num_x, dtype=torch.float64, device=dev, generator=g)
lam = torch.randn(S.num_c, dtype=torch.float64, device=dev, generator=g)
jac = torch.zeros(S.nnz_g, dtype=torch.float64, device=dev)
hes = torch.zeros(S.nnz_h, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
eng.lib.pcx_debug_read_partials.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
for r in ranges:
    lo, hi = shard_range(S.num_tiles, W, r)
    eng.set_shard(lo, hi)
    for _ in range(3):
        eng.eval_ptr(what, x, lam=lam, jac=jac, hess=hes, stream=st)
        torch.cuda.synchronize()
    buf = np.zeros(S.num_tiles * 16)
    eng.lib.pcx_debug_read_partials(eng.h, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
    t = buf.reshape(-1, 16)[lo:hi]
    t0 = t[:, 0].min()
    start, wait, node, end = [(t[:, k] - t0) / 1e3 for k in (0, 8, 2, 3)]
    dur = end - start
    smid = t[:, 10].astype(int)
    order = np.argsort(-dur)[:8]
    late = np.argsort(-end)[:6]
    print(json.dumps(dict(range=r, tiles=int(hi - lo), span_us=round(float(end.max()), 1),
                          dur_med=round(float(np.median(dur)), 2), dur_p90=round(float(np.percentile(dur, 90)), 2),
                          dur_max=round(float(dur.max()), 2),
                          node_med=round(float(np.median(node - wait)), 2), scatter_med=round(float(np.median(end - node)), 2),
                          start_p50=round(float(np.median(start)), 1), start_max=round(float(start.max()), 1),
                          slowest=[(int(i), round(float(dur[i]), 1), round(float(start[i]), 1)) for i in order],
                          last_to_end=[(int(i), round(float(start[i]), 1), round(float(end[i]), 1)) for i in late],
                          sms=int(len(set(smid.tolist()))))), flush=True)

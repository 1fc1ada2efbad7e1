"""Per-CTA phase timeline of the fused G+H kernel (debug build of the kernels)."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import build_case
from pycollo_b200 import engine as E
from examples import problems as examples
os.environ["PCX_NVRTC_EXTRA"] = "-DPCX_DEBUG_TIMELINE"
threads, tps, mb = (int(sys.argv[1]), int(sys.argv[2]) or None, int(sys.argv[3]) or None) if len(sys.argv) > 3 else (128, None, 6)
low, _, scal = build_case(examples.cart_pole_swing_up(), "lobatto", 33333, 4, seed=0, unit_scaling=True,
                          oracle=False, threads=threads, max_tile_nodes=threads, tiles_per_sm=tps)
S = low.S
eng = E.Engine(S, low.layouts, low.header, min_blocks=mb)
eng.set_scaling(*scal)
rng = np.random.default_rng(0)
R = 6
xs = [torch.from_numpy(rng.uniform(-.5, .5, S.num_x)).cuda() for _ in range(R)]
ls = [torch.from_numpy(rng.standard_normal(S.num_c)).cuda() for _ in range(R)]
js = [torch.empty(S.nnz_g, dtype=torch.float64, device="cuda") for _ in range(R)]
hs = [torch.empty(S.nnz_h, dtype=torch.float64, device="cuda") for _ in range(R)]
st = torch.cuda.current_stream().cuda_stream
what = E.EVAL_JAC | E.EVAL_HESS
for i in range(13):
    eng.eval_ptr(what, xs[i % R], lam=ls[i % R], jac=js[i % R], hess=hs[i % R], stream=st)
torch.cuda.synchronize()
buf = np.zeros(S.num_tiles * 16)
eng.lib.pcx_debug_read_partials.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
eng.lib.pcx_debug_read_partials(eng.h, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
t = buf.reshape(-1, 16)
raw_stamps = t.copy()
t0 = t[:, 0].min()
t = np.where(t == 0, np.nan, t)
rel = (t - t0) / 1e3
names = ["start", "prologue", "node", "scatter", "pre-ticket", "end"]
rel[:, 6] = np.nan_to_num(rel[:, 6], nan=-1)
print("tiles", S.num_tiles, "threads", threads, "min_blocks", mb)
for k in range(6):
    col = rel[:, k][~np.isnan(rel[:, k])]
    if col.size == 0: continue
    print(f"{names[k]:10s} min {col.min():7.2f} p10 {np.percentile(col,10):7.2f} med {np.median(col):7.2f} p90 {np.percentile(col,90):7.2f} max {col.max():7.2f} us")
rel[:, 4] = rel[:, 3]      # (the pre-ticket stamp is gone: tiles no longer take a ticket)
d = np.diff(rel[:, :6], axis=1)
for k in range(5):
    print(f"phase {names[k]}->{names[k+1]:10s}: med {np.median(d[:,k]):6.2f} p90 {np.percentile(d[:,k],90):6.2f} max {d[:,k].max():6.2f} us")
last = int(np.argmax(rel[:, 6]))
print("border CTA", last, "border end", rel[last, 6], "its pre-ticket", rel[last, 4])
print("kernel span (first start -> border end): %.2f us" % rel[last, 6])
print("prologue split: td known med %.2f, loads arrived (1st barrier) med %.2f, node map built med %.2f us"
      % (np.nanmedian(rel[:, 7]), np.nanmedian(rel[:, 8]), np.nanmedian(rel[:, 1])))
print("border CTA: tiles signalled %.2f, border end %.2f, last tile end %.2f us" % (rel[0, 9], rel[0, 6], np.nanmax(rel[:, 5])))

# write-only bandwidth ceiling
buf = torch.empty(1 << 27, dtype=torch.float64, device="cuda")   # 1 GiB
for _ in range(3): buf.fill_(1.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): buf.fill_(2.0)
e1.record(); torch.cuda.synchronize()
print("write-only (fill_) GB/s: %.0f" % (10 * buf.numel() * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9))
small = [torch.empty(35 * (1 << 20) // 8, dtype=torch.float64, device="cuda") for _ in range(6)]
for b in small: b.fill_(1.0)
torch.cuda.synchronize()
e0.record()
for i in range(60): small[i % 6].fill_(2.0)
e1.record(); torch.cuda.synchronize()
print("35 MB fill_ in ring: %.2f us each" % (1e3 * e0.elapsed_time(e1) / 60))

# per-SM view: CTAs per SM and how the node phase stretches with them
smid = raw_stamps[:, 10].astype(int)
cnt = np.bincount(smid)
print("CTAs per SM: min %d max %d (SMs used %d)" % (cnt[cnt > 0].min(), cnt.max(), (cnt > 0).sum()))
node_d = rel[:, 2] - rel[:, 1]
slow = np.argsort(-node_d)[:12]
print("slowest node phases:", [(int(i), int(smid[i]), round(float(node_d[i]), 2)) for i in slow])
per_sm_end = np.array([rel[smid == s_, 5].max() for s_ in np.flatnonzero(cnt)])
print("per-SM last CTA end: min %.2f med %.2f max %.2f us" % (per_sm_end.min(), np.median(per_sm_end), per_sm_end.max()))

# warps of one CTA: how far apart do they leave the node phase (the CTA-wide barrier
# before the scatter makes the early ones wait for the last)
w = (raw_stamps[:, 11:15] - t0) / 1e3
spread = w.max(axis=1) - w.min(axis=1)
print("intra-CTA warp skew at the end of the node phase: med %.2f p90 %.2f max %.2f us; "
      "mean wait of a warp at the barrier %.2f us" % (np.median(spread), np.percentile(spread, 90),
                                                     spread.max(), float((w.max(axis=1, keepdims=True) - w).mean())))

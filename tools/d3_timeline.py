"""Per-CTA phase durations of the fused G+H kernel on Delta III (config 4, multi-wave grid):
where a tile's lifetime goes -- table prologue, input loads, node phase, scatter.
python tools/d3_timeline.py [K]      (extra NVRTC options via PCX_NVRTC_EXTRA are kept)"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PCX_NVRTC_EXTRA"] = (os.environ.get("PCX_NVRTC_EXTRA", "") + " -DPCX_DEBUG_TIMELINE").strip()
import numpy as np, torch
from examples import problems
from examples.cases import lower_case
from pycollo_b200 import engine as E

K = int(sys.argv[1]) if len(sys.argv) > 1 else 83333
low, _, scal = lower_case(problems.delta_iii_launch_vehicle(), "lobatto", K, 4, seed=0)
S = low.S
eng = E.Engine(S, low.layouts, low.header, structure=False)
eng.set_scaling(*scal)
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
x = 0.1 + 0.3 * torch.rand(S.num_x, dtype=torch.float64, device=dev, generator=g)
lam = torch.randn(S.num_c, dtype=torch.float64, device=dev, generator=g)
jac = torch.empty(S.nnz_g, dtype=torch.float64, device=dev)
hes = torch.empty(S.nnz_h, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
what = E.EVAL_JAC | E.EVAL_HESS
for _ in range(3):
    eng.eval_ptr(what, x, lam=lam, jac=jac, hess=hes, stream=st)
torch.cuda.synchronize()
buf = np.zeros(S.num_tiles * 16)
eng.lib.pcx_debug_read_partials.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
eng.lib.pcx_debug_read_partials(eng.h, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
t = buf.reshape(-1, 16)
t = np.where(t == 0, np.nan, t)
t0 = np.nanmin(t[:, 0])
us = (t - t0) / 1e3
print("Delta III, tiles", S.num_tiles, "threads", S.threads, "opts", os.environ["PCX_NVRTC_EXTRA"])
print("kernel span %.1f us" % np.nanmax(us[:, 5]))


def stat(name, a):
    a = a[~np.isnan(a)]
    print(f"{name:34s} med {np.median(a):6.2f}  p10 {np.percentile(a,10):6.2f}  p90 {np.percentile(a,90):6.2f} us")


stat("CTA lifetime", us[:, 5] - us[:, 0])
stat("start -> tile descriptor known", us[:, 7] - us[:, 0])
stat("descriptor -> inputs in smem", us[:, 1] - us[:, 7])
stat("node phase (generated body)", us[:, 2] - us[:, 1])
if np.isnan(us[:, 4]).all():
    stat("signal + rows + scatter of G", us[:, 3] - us[:, 2])
else:                                           # two-pass form
    stat("rows + scatter of G", us[:, 4] - us[:, 2])
    stat("second pass (Hessian body)", us[:, 15] - us[:, 4])
    stat("signal + copy of H", us[:, 3] - us[:, 15])
w = us[:, 11:15]
stat("intra-CTA warp skew, node phase", np.nanmax(w, axis=1) - np.nanmin(w, axis=1))
smid = np.nan_to_num(t[:, 10]).astype(int)
cnt = np.bincount(smid)
print("CTAs per SM: min %d max %d" % (cnt[cnt > 0].min(), cnt.max()))

# lock-step diagnostic: how many CTAs start within 1 us of another CTA's start on the same SM,
# and how many tiles of one SM are in the scatter phase at the same time (time-weighted)
st, en = us[:, 0], us[:, 5]
sc0, sc1 = us[:, 2], (us[:, 4] if not np.isnan(us[:, 4]).all() else us[:, 3])
close = tot = 0
ov_num = ov_den = 0.0
for sm in np.flatnonzero(cnt):
    idx = np.flatnonzero(smid == sm)
    a = np.sort(st[idx])
    d = np.diff(a)
    near = np.zeros(a.size, bool)
    near[1:] |= d < 1.0
    near[:-1] |= d < 1.0
    close += int(near.sum()); tot += a.size
    ev = sorted([(t, 1) for t in sc0[idx]] + [(t, -1) for t in sc1[idx]])
    k = 0; last = None
    for t, dlt in ev:
        if last is not None and k > 0:
            ov_num += k * k * (t - last); ov_den += k * (t - last)
        k += dlt; last = t
print("CTAs starting within 1 us of another on the same SM: %.0f %%" % (100.0 * close / max(tot, 1)))
print("tiles of one SM scattering at the same time (seen by a scattering tile): %.2f" % (ov_num / max(ov_den, 1e-9)))
# the first CTAs of a few SMs: start / scatter begin / end (us), in start order
for sm in list(np.flatnonzero(cnt)[[0, len(np.flatnonzero(cnt)) // 2]]):
    idx = np.flatnonzero(smid == sm)
    idx = idx[np.argsort(st[idx])][:14]
    print("SM %d:" % sm, " ".join("[%.1f %.1f %.1f]" % (st[i], sc0[i], en[i]) for i in idx))
# the first generation (CTAs resident before the previous launch has completed): where does it spend its time?
g1 = st < 20.0
names = {0: "start", 7: "descriptor", 8: "after wait", 1: "inputs in smem", 2: "pass 1 done", 4: "scatter done",
         15: "pass 2 done", 3: "tile done", 5: "exit"}
print("first generation (%d CTAs), medians [p10 p90] in us:" % int(g1.sum()))
for k in (0, 7, 8, 1, 2, 4, 15, 3, 5):
    a = us[g1, k]; a = a[~np.isnan(a)]
    if a.size:
        print("  %-15s %6.1f [%6.1f %6.1f]" % (names[k], np.median(a), np.percentile(a, 10), np.percentile(a, 90)))
g2 = (st > 50.0) & (st < 65.0)
print("second generation (%d CTAs):" % int(g2.sum()))
for k in (0, 7, 8, 1, 2, 4, 15, 3, 5):
    a = us[g2, k]; a = a[~np.isnan(a)]
    if a.size:
        print("  %-15s %6.1f [%6.1f %6.1f]" % (names[k], np.median(a), np.percentile(a, 10), np.percentile(a, 90)))

"""One timing of the headline kernel under the experiment knobs given in the
environment (PCX_THREADS, PCX_TILES_PER_SM, PCX_MIN_BLOCKS, PCX_NVRTC_EXTRA, PCX_NO_PDL):
steady-state us per evaluation (200 launches from C behind a stream gate, ring of 6
buffer sets) and single-launch latency.   python tools/knobs.py [label]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from examples import problems
from examples.cases import lower_case
from pycollo_b200 import engine as E

label = sys.argv[1] if len(sys.argv) > 1 else ""
K = int(os.environ.get("KNOB_SECTIONS", 33333))
low, _, scal = lower_case(problems.cart_pole_swing_up(), "lobatto", K, 4, seed=0, unit_scaling=True)
S = low.S
eng = E.Engine(S, low.layouts, low.header, structure=False)
eng.set_scaling(*scal)
what = E.EVAL_JAC | E.EVAL_HESS
dev = torch.device("cuda")
rng = np.random.default_rng(0)
R = 6
sets = [dict(x=torch.from_numpy(rng.uniform(-0.5, 0.5, S.num_x)).to(dev),
             lam=torch.from_numpy(rng.standard_normal(S.num_c)).to(dev),
             jac=torch.empty(S.nnz_g, dtype=torch.float64, device=dev),
             hess=torch.empty(S.nnz_h, dtype=torch.float64, device=dev)) for _ in range(R)]
st = torch.cuda.current_stream().cuda_stream
args = eng.make_args(sets)
eng.eval_many(what, args, 20, stream=st, gate=False, timed=False)
torch.cuda.synchronize()
best = min(eng.eval_many(what, args, 200, stream=st, gate=True, timed=True) for _ in range(3)) / 200
short = min(eng.eval_many(what, args, 20, stream=st, gate=True, timed=True) for _ in range(3)) / 20
indep = min(eng.eval_many(what | E.EVAL_INDEPENDENT, args, 200, stream=st, gate=True, timed=True) for _ in range(3)) / 200
indep20 = min(eng.eval_many(what | E.EVAL_INDEPENDENT, args, 20, stream=st, gate=True, timed=True) for _ in range(3)) / 20
# results of the overlapped launches must equal the ordered ones
ref = [s["jac"].clone() for s in sets], [s["hess"].clone() for s in sets]
eng.eval_many(what, args, R, stream=st, gate=False, timed=False)
torch.cuda.synchronize()
same = all(torch.equal(a, s["jac"]) for a, s in zip(ref[0], sets)) and all(torch.equal(a, s["hess"]) for a, s in zip(ref[1], sets))
one = [eng.make_args([s]) for s in sets]
lat = []
for i in range(30):
    torch.cuda.synchronize()
    lat.append(eng.eval_many(what, one[i % R], 1, stream=st, gate=False, timed=True))
alg = 8 * (S.num_x + S.nnz_g) + 8 * (S.num_x + S.num_c + S.nnz_h)
print(json.dumps(dict(label=label, tiles=int(S.num_tiles), threads=int(S.threads),
                      us_per_eval_200=round(1e3 * best, 3), us_per_eval_20=round(1e3 * short, 3),
                      frac_200=round(alg / (best * 1e-3) / 6553e9, 4),
                      frac_20=round(alg / (short * 1e-3) / 6553e9, 4),
                      latency_us=round(1e3 * float(np.median(lat[5:])), 3),
                      us_independent_200=round(1e3 * indep, 3), us_independent_20=round(1e3 * indep20, 3),
                      frac_independent_20=round(alg / (indep20 * 1e-3) / 6553e9, 4), independent_equal=bool(same),
                      env={k: v for k, v in os.environ.items() if k.startswith("PCX_")})), flush=True)

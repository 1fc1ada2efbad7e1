"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every callback
of three small problems incl. a multi-phase one, the mesh-error pipeline and a
sharded two-stage evaluation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import build_case, make_engine
from pycollo_b200 import engine as E
from examples import problems as examples
from pycollo_b200.parallel import shard_range

ALL = E.EVAL_C | E.EVAL_DY | E.EVAL_JAC | E.EVAL_HESS | E.EVAL_F | E.EVAL_GRAD
rng = np.random.default_rng(0)
for name, K, nodes, kw in (("cart_pole_swing_up", 40, 4, {}), ("double_pendulum", 6, [4, 7, 2, 10, 3, 5], dict(max_tile_nodes=20)),
                           ("multiphase_sliding_mass", 30, 4, dict(max_tile_nodes=24))):
    low, _, scal = build_case(getattr(examples, name)(), "lobatto", K, nodes, oracle=False, **kw)
    eng = make_engine(low, scal)
    x, lam = rng.uniform(-0.5, 0.5, low.S.num_x), rng.standard_normal(low.S.num_c)
    out = eng.eval_host(ALL, x, lam, 0.7)
    eng.eval_host(E.EVAL_JAC | E.EVAL_HESS, x, lam, 0.7)
    assert all(np.all(np.isfinite(v)) for v in out.values()), name
    x_ph = eng.refit_to_ph_host(x, out["dy"][0])
    eng.set_shard(*shard_range(low.S.num_tiles, 2, 0))
    eng.eval_host(E.EVAL_JAC, x)
    print(name, "ok", low.S.num_tiles, "tiles", flush=True)

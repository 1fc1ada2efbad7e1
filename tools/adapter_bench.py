"""cyipopt-style callback adapter (pycollo_b200/nlp.py) at BASELINE config 2: the
five callbacks of one solver iterate -- objective, gradient, constraints, jacobian,
hessian at a NEW x -- host numpy in / host numpy out, row-major / lower-triangular
ordering: what an IPOPT host pays per iterate.  Variants: how the object decides
that x is unchanged (exact compare, sampled compare, the new_x flag of IPOPT's TNLP
interface)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from examples import problems as examples
from pycollo_b200.nlp import NlpCallbacks

ocp = examples.cart_pole_swing_up()
ocp.settings.scaling_method = "none"
examples.set_mesh(ocp, 33333, 4)
ocp.initialise()
it = ocp._backend.current_iteration
rng = np.random.default_rng(0)
xs = [rng.uniform(-0.5, 0.5, it.num_x) for _ in range(4)]
lam = rng.standard_normal(it.num_c)


def iterate(cb, x, flag):
    kw = (lambda first: {}) if flag is None else (lambda first: dict(new_x=first))
    cb.objective(x, **kw(True)); cb.gradient(x, **kw(False)); cb.constraints(x, **kw(False))
    cb.jacobian(x, **kw(False)); cb.hessian(x, lam, 1.0, **kw(False))


for ordering, x_check, flag, reg in (("cyipopt", "full", None, True), ("cyipopt", "sampled", None, True),
                                     ("cyipopt", "full", True, False), ("cyipopt", "full", True, True),
                                     ("casadi", "full", True, True)):
    cb = NlpCallbacks(it, ordering, x_check, register_inputs=reg)
    for k in range(3):
        iterate(cb, xs[k % 4], flag)
    n = 30
    t0 = time.perf_counter()
    for k in range(n):
        iterate(cb, xs[k % 4], flag)
    dt = (time.perf_counter() - t0) / n
    print(json.dumps(dict(adapter="NlpCallbacks", ordering=ordering,
                          same_x_test=("new_x flag" if flag else x_check), nodes=100000,
                          upload=("DMA from the caller's arrays (registered once)"
                                  if cb.num_registered_uploads else "staging copy"),
                          callbacks_per_iterate=5, x_uploads_per_iterate=cb.num_x_uploads / (n + 3),
                          ms_per_iterate=round(1e3 * dt, 3), iterates_per_s=round(1 / dt, 1))), flush=True)
    cb.close()

"""cyipopt-style callback adapter (pycollo_b200/nlp.py) at BASELINE config 2:
jacobian(x) + hessian(x, lam, sigma) per "iteration", host numpy in / host numpy
out, row-major / lower-triangular ordering -- what an IPOPT host pays per iterate."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from examples import problems as examples
from pycollo_b200.backend import Cuda
from pycollo_b200.nlp import NlpCallbacks

ocp = examples.cart_pole_swing_up()
ocp.settings.scaling_method = "none"
examples.set_mesh(ocp, 33333, 4)
backend = Cuda(ocp)
for step in ("create_bounds", "create_scaling", "create_quadrature",
             "create_initial_mesh", "create_guess", "create_mesh_iterations"):
    getattr(backend, step)()
it = backend.current_iteration
it.generate_nlp()
rng = np.random.default_rng(0)
x = rng.uniform(-0.5, 0.5, it.num_x)
lam = rng.standard_normal(it.num_c)
for ordering in ("cyipopt", "casadi"):
    cb = NlpCallbacks(it, ordering)
    for _ in range(3):
        cb.jacobian(x); cb.hessian(x, lam, 1.0)
    n = 30
    t0 = time.perf_counter()
    for _ in range(n):
        cb.jacobian(x); cb.hessian(x, lam, 1.0)
    dt = (time.perf_counter() - t0) / n
    print(json.dumps(dict(adapter="NlpCallbacks", ordering=ordering, nodes=100000,
                          ms_per_jac_plus_hess=round(1e3 * dt, 3), evals_per_s=round(1 / dt, 1))), flush=True)

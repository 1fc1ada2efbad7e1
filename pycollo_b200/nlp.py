"""cyipopt-style callback object over the CUDA engine.

The method set and argument order are those of the reference's (dead but
explicit) IPOPT plumbing: ``IPOPTProblem`` in ``pycollo/nlp.py:36-76`` --
``objective(x)``, ``gradient(x)``, ``constraints(x)``, ``jacobian(x)``,
``jacobianstructure()``, ``hessian(x, lagrange, obj_factor)``,
``hessianstructure()``, ``intermediate(...)``.  That path orders the Jacobian
row-major and uses the *lower* triangle of the Hessian
(``pycollo/iteration.py:930-933, 965-968, 1057``), whereas the engine stores the
live backend's CasADi order (CCS / upper triangle).  The two fixed permutations
are computed once here.
"""
from __future__ import annotations

import numpy as np

from . import engine as _engine


class NlpCallbacks:
    def __init__(self, iteration, ordering="cyipopt"):
        self.it = iteration
        S = iteration.S
        gr, gc = S.G_structure()
        hr, hc = S.H_structure()
        if ordering == "cyipopt":
            self.g_perm = np.lexsort((gc, gr))             # row-major
            self.h_perm = np.lexsort((hr, hc))             # tril, row-major
            self._g_struct = (gr[self.g_perm], gc[self.g_perm])
            # lower triangle: swap (row, col) of the stored upper triangle
            self._h_struct = (hc[self.h_perm], hr[self.h_perm])
        elif ordering == "casadi":
            self.g_perm = np.arange(len(gr))
            self.h_perm = np.arange(len(hr))
            self._g_struct = (gr, gc)
            self._h_struct = (hr, hc)
        else:
            raise ValueError("ordering must be 'cyipopt' or 'casadi'")
        self.num_evals = dict(objective=0, gradient=0, constraints=0,
                              jacobian=0, hessian=0)

    def objective(self, x):
        self.num_evals["objective"] += 1
        return float(self.it.evaluate(_engine.EVAL_F, x)["f"][0])

    def gradient(self, x):
        self.num_evals["gradient"] += 1
        return self.it.evaluate(_engine.EVAL_GRAD, x)["grad"][0]

    def constraints(self, x):
        self.num_evals["constraints"] += 1
        return self.it.evaluate(_engine.EVAL_C, x)["c"][0]

    def jacobian(self, x):
        self.num_evals["jacobian"] += 1
        return self.it.evaluate(_engine.EVAL_JAC, x)["jac"][0][self.g_perm]

    def jacobianstructure(self):
        return self._g_struct

    def hessian(self, x, lagrange, obj_factor):
        self.num_evals["hessian"] += 1
        return self.it.evaluate(_engine.EVAL_HESS, x, lagrange,
                                obj_factor)["hess"][0][self.h_perm]

    def hessianstructure(self):
        return self._h_struct

    def intermediate(self, alg_mod, iter_count, obj_value, inf_pr, inf_du, mu,
                     d_norm, regularization_size, alpha_du, alpha_pr, ls_trials):
        self.last_iterate = dict(iter_count=iter_count, obj_value=obj_value,
                                 inf_pr=inf_pr, inf_du=inf_du)

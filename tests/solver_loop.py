"""A host NLP solver loop over cyipopt-style callback objects (TEST SIDE).

IPOPT / cyipopt / CasADi are absent from this image, so the reference's
``solve_nlp`` (``pycollo/backend.py:1807-1827``) cannot run; the closest stand-in
that consumes exactly the same callback contract (``pycollo/nlp.py:36-76``:
objective, gradient, constraints, jacobian + structure, hessian + structure) is
the package's built-in interior-point Newton method (``pycollo_b200/ipnewton.py``, on
scipy's sparse LU; what ``Cuda.solve_nlp`` runs when cyipopt is absent).  It is used to check that G and H are JOINTLY right (a solve converges
to the reference's pinned objectives) and that the CUDA callbacks and the CPU
oracle drive the solver through the SAME iterates (identical iteration counts).
"""
import numpy as np
import scipy.sparse as sp


class OracleCallbacks:
    """The callback contract served by the CPU oracle (``oracle/blockwise.py``) in
    the cyipopt ordering: row-major Jacobian, lower-triangular Hessian."""

    def __init__(self, B):
        self.B = B
        gr, gc = B.G_structure()
        self.g_perm = np.lexsort((gc, gr))
        self._g = (gr[self.g_perm], gc[self.g_perm])
        hr, hc = B.H_structure()
        self._h = (hc, hr)

    def objective(self, x):
        return float(self.B.J(x))

    def gradient(self, x):
        return np.asarray(self.B.g(x), dtype=float)

    def constraints(self, x):
        return np.asarray(self.B.c(x), dtype=float)

    def jacobian(self, x):
        return self.B.G_nonzeros(x)[self.g_perm]

    def jacobianstructure(self):
        return self._g

    def hessian(self, x, lagrange, obj_factor):
        return self.B.H_nonzeros(x, obj_factor, lagrange)

    def hessianstructure(self):
        return self._h


from pycollo_b200.ipnewton import solve  # noqa: E402,F401  (the solver the tests drive)

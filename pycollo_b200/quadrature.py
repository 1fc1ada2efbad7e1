"""Collocation point sets, quadrature weights and integration blocks.

Host-side, once-per-settings tables that end up as device-resident constants
of the CUDA engine.

Two constructions are offered (``Quadrature(method, tables=...)``):

``tables="reference"`` (default) reproduces the NUMERICS of the reference's
generator (``pycollo/quadrature.py:116-261``): points from numpy's Legendre
companion-matrix roots, the closed-form weights, and the Butcher rows as the
``numpy.linalg.solve`` solution of the simplifying-condition system
``sum_i w_i c_i^k a_ij = w_j (1 - c_j^(k+1))/(k+1) - w_last w_j`` assembled in the
same unknown/equation order -- so the tables are the reference's to the last
bit (``tests/test_quadrature_mesh.py`` against tables produced by executing the
reference, orders 2..20), including the digits that ill-conditioned solve loses
at high order (1.3e-12 at order 10).  Parity is with the reference, not with
the mathematics.

``tables="exact"`` is written from the mathematical definitions instead:

* Lobatto (LGL) points are the roots of ``P'_{n-1}`` plus the two ends, found by
  Newton iteration on the Legendre recurrence; weights ``1/(n(n-1)P_{n-1}(x)^2)``
  (the reference normalises Lobatto weights to sum to ONE, see
  ``pycollo/quadrature.py:201-203`` and ``tests/unit/test_quadrature.py:48-55``).
* Radau (LGR, left end included) points are the ``n-1`` roots of
  ``P_{n-2}+P_{n-1}`` and a padding entry; weights sum to TWO and the padding
  weight is exactly 0 (``pycollo/quadrature.py:116-139``).  This asymmetry
  (sum 1 vs sum 2) is the reference's behaviour and is kept on purpose
  ("Radau quirk").
* The integration block ``A`` (``(n-1) x n``) is the collocation matrix
  ``A[l, j] = integral_0^{c_{l+1}} L_j(t) dt`` on the unit interval, which is the
  unique solution of the simplifying conditions the reference solves
  numerically (``pycollo/quadrature.py:141-163, 214-246``).  Its last row is
  the weight vector on [0, 1].  Under Radau the last column is exactly zero.

Agreement with the reference's own tables is pinned by
``tests/test_quadrature_mesh.py`` against ``tests/golden/quadrature_*.npz``.
"""
from __future__ import annotations

import numpy as np

LOBATTO = "lobatto"
RADAU = "radau"
GAUSS = "gauss"
QUADRATURE_METHODS = (LOBATTO, RADAU)
DEFAULT_COLLOCATION_POINTS_MIN = 4
DEFAULT_COLLOCATION_POINTS_MAX = 10


def _legendre_pair(n, x):
    """Return (P_n(x), P_{n-1}(x)) by the three-term recurrence (vectorised)."""
    x = np.asarray(x, dtype=np.float64)
    p_prev = np.ones_like(x)
    if n == 0:
        return p_prev, np.zeros_like(x)
    p = x.copy()
    for k in range(1, n):
        p, p_prev = ((2 * k + 1) * x * p - k * p_prev) / (k + 1), p
    return p, p_prev


def _legendre_derivs(n, x):
    """P_n, P_n', P_n'' at interior points x (|x| < 1)."""
    p, pm1 = _legendre_pair(n, x)
    one_m_x2 = 1.0 - x * x
    dp = n * (pm1 - x * p) / one_m_x2
    ddp = (2.0 * x * dp - n * (n + 1) * p) / one_m_x2
    return p, dp, ddp


def _lobatto_points(n):
    if n == 2:
        return np.array([-1.0, 1.0])
    m = n - 1
    k = np.arange(1, m)
    x = -np.cos(np.pi * k / m)  # Chebyshev-Lobatto start
    for _ in range(100):
        _, dp, ddp = _legendre_derivs(m, x)
        step = dp / ddp
        x = x - step
        if np.max(np.abs(step)) < 1e-16:
            break
    x = 0.5 * (x - x[::-1])  # enforce exact antisymmetry
    return np.concatenate([[-1.0], x, [1.0]])


def _radau_points(n):
    """The n-1 left-Radau points (roots of P_{n-2} + P_{n-1}), ascending."""
    m = n - 1  # number of true points
    if m == 1:
        return np.array([-1.0])
    # interior roots of (P_{m-1}+P_m)/(1+x); start from Chebyshev-like guesses
    k = np.arange(1, m)
    x = -np.cos(2.0 * np.pi * k / (2 * m - 1))
    for _ in range(200):
        pm, pm1 = _legendre_pair(m, x)
        f = pm + pm1
        one_m_x2 = 1.0 - x * x
        dpm = m * (pm1 - x * pm) / one_m_x2
        if m - 1 == 0:
            dpm1 = np.zeros_like(x)
        else:
            pm1_, pm2 = _legendre_pair(m - 1, x)
            dpm1 = (m - 1) * (pm2 - x * pm1_) / one_m_x2
        df = dpm + dpm1
        # deflate the known root at -1: g = f/(1+x)
        g = f / (1.0 + x)
        dg = (df - g) / (1.0 + x)
        step = g / dg
        x = x - step
        if np.max(np.abs(step)) < 1e-16:
            break
    return np.concatenate([[-1.0], np.sort(x)])


def _collocation_matrix(c):
    """A[i, j] = int_0^{c_i} L_j(t) dt for distinct abscissae c on [0, 1]."""
    c = np.asarray(c, dtype=np.float64)
    s = len(c)
    # barycentric weights
    diff = c[:, None] - c[None, :]
    np.fill_diagonal(diff, 1.0)
    bw = 1.0 / np.prod(diff, axis=1)
    ng = s // 2 + 2
    xg, wg = np.polynomial.legendre.leggauss(ng)
    A = np.zeros((s, s))
    for i in range(s):
        if c[i] == 0.0:
            continue
        t = 0.5 * c[i] * (xg + 1.0)
        w = 0.5 * c[i] * wg
        # Lagrange basis at quadrature abscissae via the product form
        L = np.ones((len(t), s))
        for j in range(s):
            for k in range(s):
                if k != j:
                    L[:, j] *= (t - c[k]) / (c[j] - c[k])
        A[i, :] = w @ L
    del bw
    return A


def _reference_numerics(n, method):
    """Points, weights and Butcher array with the floating-point operations of
    the reference's generators (``pycollo/quadrature.py:116-163`` Radau,
    ``:189-246`` Lobatto), so that every entry is bit-identical to what the
    reference computes on the same numpy/LAPACK.

    The Butcher rows 1..n-2 solve, for k = 0..n-3 and every column j,
        sum_{i=1}^{n-2} w_i c_i^k a_ij = (w_j/(k+1)) (1 - c_j^(k+1)) - w_{n-1} w_j
    (c = points on [0, 1]).  The reference assembles all columns into ONE
    n(n-2)-square system (equation (k, j) -> row j + k n, unknown a_ij -> column
    (i-1) + j (n-2)) and calls ``numpy.linalg.solve``; the same matrix is formed
    here (a blocked LU of the single system and n small solves need not agree
    in the last bit, and the conditioning amplifies any such difference)."""
    leg = np.polynomial.legendre.Legendre
    if method == LOBATTO:
        poly = leg([0] * (n - 1) + [1])
        x = np.append(np.insert(poly.deriv().roots(), 0, -1, axis=0), 1)
        w = np.array([1 / (n * (n - 1) * (poly(v) ** 2)) for v in x])
        last_row = w
    else:
        x = np.concatenate([leg([0] * (n - 2) + [1, 1]).roots(), np.array([0])])
        lower = leg([0] * (n - 2) + [1])
        w = np.array([2 / (n - 1) ** 2]
                     + [(1 - v) / ((n - 1) ** 2 * (lower(v) ** 2)) for v in x[1:-1]])
        w = np.concatenate([w, np.array([0])])
        last_row = w / 2
    c = 0.5 * x + 0.5
    butcher = np.zeros((n, n))
    butcher[-1, :] = last_row
    m = n - 2
    if m > 0:
        M = np.zeros((n * m, n * m))
        rhs = np.zeros(n * m)
        for k in range(m):
            coef = [w[i + 1] * c[i + 1] ** k for i in range(m)]
            for j in range(n):
                M[j + k * n, j * m:(j + 1) * m] = coef
                rhs[j + k * n] = (w[j] / (k + 1)) * (1 - c[j] ** (k + 1)) - w[-1] * w[j]
        butcher[1:-1, :] = np.linalg.solve(M, rhs).reshape(m, -1, order="F")
    return {"points": x, "weights": w, "butcher": butcher}


class Quadrature:
    """Tables for one quadrature scheme; orders are generated lazily.

    Mirrors the query surface of the reference class
    (``pycollo/quadrature.py:40-114``): ``quadrature_point(order, domain=)``,
    ``quadrature_weight(order)``, ``butcher_array(order)``, ``A_matrix(order)``
    (integration block, rows 1.. of the Butcher array) and ``D_matrix(order)``
    (the ``[1 | -I]`` difference block).
    """

    def __init__(self, method=LOBATTO, tables="reference"):
        if method == GAUSS:
            raise ValueError("gauss quadrature is unsupported "
                             "(pycollo/quadrature.py:34-35)")
        if method not in QUADRATURE_METHODS:
            raise ValueError(f"unknown quadrature method {method!r}")
        if tables not in ("reference", "exact"):
            raise ValueError("tables must be 'reference' or 'exact'")
        self.method = method
        self.tables = tables
        self._cache = {}

    @classmethod
    def adopt(cls, source, method=None):
        """Drop-in path: take points / weights / Butcher arrays from a quadrature
        object produced elsewhere -- the reference's own ``Quadrature``
        (``pycollo/quadrature.py:40-114``: ``quadrature_point(order)``,
        ``quadrature_weight(order)``, ``butcher_array(order)``) or a mapping with
        ``points_<n>``, ``weights_<n>``, ``butcher_<n>`` entries -- instead of
        generating them, so that the integration blocks ``A(N_k) * h_k`` are the
        reference's to the last bit (including the digits its ill-conditioned
        solve loses at high order)."""
        if method is None:
            method = getattr(source, "method", None) or \
                source.backend.ocp.settings.quadrature_method
        self = cls(method)
        self._source = source
        return self

    def _adopted(self, order):
        src = self._source
        if hasattr(src, "butcher_array"):
            return {"points": np.asarray(src.quadrature_point(order), dtype=np.float64),
                    "weights": np.asarray(src.quadrature_weight(order), dtype=np.float64),
                    "butcher": np.asarray(src.butcher_array(order), dtype=np.float64)}
        return {"points": np.asarray(src[f"points_{order}"], dtype=np.float64),
                "weights": np.asarray(src[f"weights_{order}"], dtype=np.float64),
                "butcher": np.asarray(src[f"butcher_{order}"], dtype=np.float64)}

    def _tables(self, order):
        order = int(order)
        if order < 2:
            raise ValueError("a mesh section needs at least two nodes")
        tab = self._cache.get(order)
        if tab is None and getattr(self, "_source", None) is not None:
            try:
                tab = self._cache[order] = self._adopted(order)
            except KeyError:            # order absent from an adopted table set:
                tab = None              # generate it (only auxiliary operators ask)
        if tab is None:
            if self.tables == "reference":
                tab = _reference_numerics(order, self.method)
            else:
                tab = (self._lobatto(order) if self.method == LOBATTO
                       else self._radau(order))
            self._cache[order] = tab
        return tab

    @staticmethod
    def _lobatto(n):
        x = _lobatto_points(n)
        p, _ = _legendre_pair(n - 1, x)
        w = 1.0 / (n * (n - 1) * p * p)
        c = 0.5 * (x + 1.0)
        butcher = _collocation_matrix(c)
        butcher[0, :] = 0.0
        butcher[-1, :] = w
        return {"points": x, "weights": w, "butcher": butcher}

    @staticmethod
    def _radau(n):
        xr = _radau_points(n)
        m = n - 1
        pm1, _ = _legendre_pair(m - 1, xr) if m >= 1 else (np.ones_like(xr), None)
        w_true = np.empty(m)
        w_true[0] = 2.0 / m ** 2
        if m > 1:
            w_true[1:] = (1.0 - xr[1:]) / (m ** 2 * pm1[1:] ** 2)
        x = np.concatenate([xr, [0.0]])
        w = np.concatenate([w_true, [0.0]])
        c = 0.5 * (xr + 1.0)
        inner = _collocation_matrix(c)          # (n-1) x (n-1), row 0 is zero
        butcher = np.zeros((n, n))
        butcher[:m, :m] = inner
        butcher[0, :] = 0.0
        butcher[-1, :] = w / 2.0
        return {"points": x, "weights": w, "butcher": butcher}

    def quadrature_point(self, order, *, domain=None):
        pts = self._tables(order)["points"]
        if domain is not None:
            stretch = 0.5 * (domain[1] - domain[0])
            shift = 0.5 * (domain[0] + domain[1])
            return stretch * pts + shift
        return pts

    def quadrature_weight(self, order):
        return self._tables(order)["weights"]

    def butcher_array(self, order):
        return self._tables(order)["butcher"]

    def A_matrix(self, order):
        """Integration block: (order-1) x order (``pycollo/quadrature.py:170``)."""
        return self._tables(order)["butcher"][1:, :]

    def refit_matrices(self, order):
        """Per-section re-fit operators of the solution post-processing
        (``pycollo/solution/solution_abc.py:60-142``) in matrix form, on the local
        coordinate ``xi`` in [-1, 1] of a section with ``order`` nodes, evaluated
        at the ``order - 1`` interior nodes ``zeta_j`` of the same section on the
        p+1 mesh (``order + 1`` nodes, ``mesh_refinement.py:75-86``):

        * ``Cy[j, m] = int_{-1}^{zeta_j} L_m(xi) dxi`` -- the reference fits a
          Legendre series through ``dy * T/2`` at the section's nodes and
          integrates it from the section start (``Legendre.fit(...).integ(k=y0)``):
          ``y_ph[j] = y[start] + stretch * (h_k / 2) * sum_m Cy[j, m] dy[m]``.
          Lobatto interpolates all ``order`` nodes (``:86-91``); Radau the first
          ``order - 1`` (``:117-123``), so its last column is zero.
        * ``Pu[j, m] = L_m(zeta_j)`` over all ``order`` nodes
          (``Polynomial.fit(t_k, u_k, deg=order-1)``, ``:94-99, 135-140``).
        """
        n = int(order)
        key = ("refit", n)
        hit = self._cache.get(key)
        if hit is not None:
            return hit
        leg = np.polynomial.legendre
        xi = np.array(self.quadrature_point(n), dtype=np.float64)
        zeta = np.array(self.quadrature_point(n + 1), dtype=np.float64)
        if self.method == RADAU:                 # the stored last point is a placeholder
            xi[-1] = 1.0
            zeta[-1] = 1.0
        zin = zeta[1:n]

        def basis(nodes):
            V = leg.legvander(nodes, len(nodes) - 1)
            return np.linalg.solve(V, np.eye(len(nodes)))      # column m: series of L_m

        Pu = leg.legval(zin, basis(xi)).T                       # (n-1, n)
        fit_nodes = xi if self.method == LOBATTO else xi[:-1]
        coef = basis(fit_nodes)
        Cy = np.zeros((n - 1, n))
        for m in range(len(fit_nodes)):
            integ = leg.legint(coef[:, m], lbnd=-1.0)
            Cy[:, m] = leg.legval(zin, integ)
        self._cache[key] = (Cy, Pu)
        return Cy, Pu

    def D_matrix(self, order):
        """Difference block ``[1 | -I]`` (``pycollo/quadrature.py:165-168``)."""
        n = int(order)
        return np.hstack([np.ones((n - 1, 1)), -np.eye(n - 1)])

"""ORACLE (1/2) -- literal symbolic expansion of the reference NLP.  TEST ONLY.

Follows ``pycollo/backend.py:1433-1679`` statement by statement, with sympy in
place of CasADi SX:

* ``create_iteration_specific_variable_symbols`` ``:1433-1457`` -- one symbol per
  (variable, node); endpoint symbols are node 0 / node N-1 ``:1439-1442``;
  ``x = [phase: y.., u.., q.., t..] + [s..]``, variable-major, node-minor.
* every OCP symbol is the scaled expression ``V * x_tilde + r`` ``:1459-1463``
  (``backend.py:170-187``).
* ``expand_eqn_to_vec`` ``:1565-1570`` -- an equation is substituted once per node.
* defect ``:1572-1603``:  ``W_d * (A @ y + 0.5 * (tF - t0) * I @ f)`` with
  ``A = mesh.sA_matrix`` (differences) and ``I = mesh.sI_matrix`` (integration).
* path ``:1605-1616``, integral ``:1618-1647`` (``q - 0.5 (tF - t0) W . g``),
  endpoint ``:1649-1655``; c ordering ``:1551-1563``.
* ``J = w * J`` ``:1495-1504``; ``g = gradient`` ``:1506-1511``;
  ``G = jacobian(c, x)`` ``:1674-1679`` reported in CCS order like
  ``evaluate_G_structure`` ``:1747-1761``; H = upper triangle (CCS) of the
  Hessian of ``sigma * J + lam . c`` (CasADi ``nlpsol`` convention, ``:1693``).

Structural nonzeros are the entries whose *symbolic* derivative is not the
zero expression.  Exact-zero mesh coefficients vanish on multiplication (sympy
folds ``0.0 * expr`` like CasADi SX folds ``0 * x``), which gives the "pruned"
pattern of SURVEY.md §7; pass ``prune=False`` to keep them structural.

Only for small meshes: cost grows with N * expression size.
"""
from __future__ import annotations

import numpy as np
import sympy as sym

from .common import lower, ocp_variable_offsets


class ExpandedNLP:
    def __init__(self, ocp, bounds, meshes, *, scaling_method="bounds",
                 w=1.0, W_ocp=None, prune=True, want_hessian=True):
        """``meshes``: per phase dict(N, sA (scipy), sI (scipy), W (array))."""
        lp = lower(ocp, bounds, scaling_method)
        self.lp = lp
        offs, n_ocp = ocp_variable_offsets(lp)
        V, r = lp.V_ocp, lp.r_ocp
        n_s = len(lp.s)
        s_off = n_ocp

        # ---- iteration symbols (backend.py:1433-1457) ----
        x_syms = []
        node_maps = []       # per phase: list over nodes of {user sym: V*xt + r}
        point_map = {}
        time_expr = []
        for ip, (ph, off, mesh) in enumerate(zip(lp.phases, offs, meshes)):
            N = int(mesh["N"])
            cols = {}
            for kind, syms_, o in (("y", ph.y, off["y"]), ("u", ph.u, off["u"])):
                for j, v in enumerate(syms_):
                    xt = sym.symbols(f"_{kind}{j}_P{ip}_n0:{N}")
                    x_syms.extend(xt)
                    cols[v] = [V[o + j] * a + r[o + j] for a in xt]
            for j, v in enumerate(ph.q):
                xt = sym.Symbol(f"_q{j}_P{ip}")
                x_syms.append(xt)
                point_map[v] = V[off["q"] + j] * xt + r[off["q"] + j]
            for j, v in enumerate(ph.t):
                xt = sym.Symbol(f"_t{j}_P{ip}")
                x_syms.append(xt)
                point_map[v] = V[off["t"] + j] * xt + r[off["t"] + j]
            for v, a, c in zip(ph.y, ph.y_t0, ph.y_tF):
                point_map[a] = cols[v][0]
                point_map[c] = cols[v][-1]
            node_maps.append([{v: cols[v][m] for v in cols} for m in range(N)])
            time_expr.append(None)
        for j, v in enumerate(lp.s):
            xt = sym.Symbol(f"_s{j}")
            x_syms.append(xt)
            point_map[v] = V[s_off + j] * xt + r[s_off + j]
        s_map = {v: point_map[v] for v in lp.s}
        self.x_syms = x_syms
        self.num_x = len(x_syms)

        # ---- constraint scaling per OCP-level constraint (backend.py:1465-1493)
        n_c_ocp = sum(len(ph.f) + len(ph.p) + len(ph.g) for ph in lp.phases) \
            + len(lp.b)
        W_ocp = np.ones(n_c_ocp) if W_ocp is None else np.asarray(W_ocp, float)
        assert W_ocp.shape == (n_c_ocp,)

        # ---- c (backend.py:1513-1672) ----
        c = []
        dy = []
        kW = 0
        for ip, (ph, mesh) in enumerate(zip(lp.phases, meshes)):
            N = int(mesh["N"])
            A = mesh["sA"].tocoo()
            I = mesh["sI"].tocoo()
            Wq = np.asarray(mesh["W"], dtype=float)
            t0 = ph.t0.subs(point_map) if isinstance(ph.t0, sym.Symbol) else ph.t0
            tF = ph.tF.subs(point_map) if isinstance(ph.tF, sym.Symbol) else ph.tF
            half_dt = 0.5 * (tF - t0)
            nm = node_maps[ip]

            def expand(e, nm=nm):
                return [e.subs({**mp, **s_map}, simultaneous=True) for mp in nm]

            for i, (v, fe) in enumerate(zip(ph.y, ph.f)):
                yv = [nm[m][v] for m in range(N)]
                fv = expand(fe)
                dy.extend(fv)
                Ay = [0] * (N - 1)
                If = [0] * (N - 1)
                for rr, cc, val in zip(A.row, A.col, A.data):
                    Ay[rr] = Ay[rr] + float(val) * yv[cc]
                for rr, cc, val in zip(I.row, I.col, I.data):
                    coef = float(val)
                    if coef == 0.0 and not prune:
                        coef = sym.Symbol("_ZERO_")   # keeps the dependency
                    If[rr] = If[rr] + coef * fv[cc]
                Wd = float(W_ocp[kW])
                kW += 1
                c.extend(Wd * (a + half_dt * b) for a, b in zip(Ay, If))
            for pe in ph.p:
                Wp = float(W_ocp[kW])
                kW += 1
                c.extend(Wp * e for e in expand(pe))
            for i, (qv, ge) in enumerate(zip(ph.q, ph.g)):
                Wi = float(W_ocp[kW])
                kW += 1
                gv = expand(ge)
                acc = 0
                for m in range(N):
                    coef = float(Wq[m])
                    if coef == 0.0 and not prune:
                        coef = sym.Symbol("_ZERO_")
                    acc = acc + coef * gv[m]
                c.append(Wi * (point_map[qv] - half_dt * acc))
        for be in lp.b:
            We = float(W_ocp[kW])
            kW += 1
            c.append(We * be.subs(point_map, simultaneous=True))
        zero_sub = {sym.Symbol("_ZERO_"): 0.0}
        self.c_exprs = [sym.sympify(e) for e in c]
        self.dy_exprs = [sym.sympify(e) for e in dy]
        self.num_c = len(c)
        self.J_expr = float(w) * lp.J.subs(point_map, simultaneous=True)

        xs = self.x_syms
        index = {s: i for i, s in enumerate(xs)}
        # ---- g ----
        g_exprs = [sym.diff(self.J_expr, s) for s in xs]
        # ---- G in CCS order (col-major), structural = symbolic non-zero ----
        entries = []
        for ri, e in enumerate(self.c_exprs):
            for s in sorted(e.free_symbols & set(xs), key=index.get):
                d = sym.diff(e, s)
                if d != 0:
                    entries.append((index[s], ri, d))
        entries.sort(key=lambda t: (t[0], t[1]))
        self.G_cols = np.array([t[0] for t in entries], dtype=np.int64)
        self.G_rows = np.array([t[1] for t in entries], dtype=np.int64)
        G_exprs = [t[2].subs(zero_sub) for t in entries]

        # ---- H: triu(CCS) of Hessian of sigma*J + lam.c ----
        self.H_rows = self.H_cols = None
        H_exprs = []
        if want_hessian:
            sigma = sym.Symbol("_sigma_")
            lam = sym.symbols(f"_lam_0:{self.num_c}") if self.num_c else ()
            hess = {}

            def add_hessian(e, mult):
                fs = sorted(e.free_symbols & set(xs), key=index.get)
                for ia, a in enumerate(fs):
                    da = sym.diff(e, a)
                    if da == 0:
                        continue
                    for b_ in fs[ia:]:
                        d2 = sym.diff(da, b_)
                        if d2 != 0:
                            key = (index[b_], index[a])     # (col, row), row<=col
                            hess[key] = hess.get(key, 0) + mult * d2

            add_hessian(self.J_expr, sigma)
            for e, l in zip(self.c_exprs, lam):
                add_hessian(e, l)
            keys = sorted(hess)
            self.H_cols = np.array([k[0] for k in keys], dtype=np.int64)
            self.H_rows = np.array([k[1] for k in keys], dtype=np.int64)
            H_exprs = [hess[k].subs(zero_sub) for k in keys]
            self._H_fn = sym.lambdify([xs, sigma, list(lam)], H_exprs,
                                      modules="numpy", cse=True)
        self.c_exprs = [e.subs(zero_sub) for e in self.c_exprs]
        self._J_fn = sym.lambdify([xs], self.J_expr, modules="numpy")
        self._g_fn = sym.lambdify([xs], g_exprs, modules="numpy", cse=True)
        self._c_fn = sym.lambdify([xs], self.c_exprs, modules="numpy", cse=True)
        self._dy_fn = sym.lambdify([xs], self.dy_exprs, modules="numpy", cse=True)
        self._G_fn = sym.lambdify([xs], G_exprs, modules="numpy", cse=True)

    # ---- callbacks (backend.py:1713-1771) ----
    def J(self, x):
        return float(self._J_fn(list(x)))

    def g(self, x):
        return np.array(self._g_fn(list(x)), dtype=float).reshape(self.num_x)

    def c(self, x):
        return np.array(self._c_fn(list(x)), dtype=float).reshape(self.num_c)

    def dy(self, x):
        return np.array(self._dy_fn(list(x)), dtype=float).ravel()

    def G_structure(self):
        return self.G_rows, self.G_cols

    def G_nonzeros(self, x):
        return np.array(self._G_fn(list(x)), dtype=float).reshape(len(self.G_rows))

    def H_structure(self):
        return self.H_rows, self.H_cols

    def H_nonzeros(self, x, sigma, lam):
        return np.array(self._H_fn(list(x), float(sigma), list(lam)),
                        dtype=float).reshape(len(self.H_rows))

"""ORACLE (2/2) -- vectorised numpy restatement of the reference callbacks.  TEST ONLY.

Same mathematics as ``oracle/expand.py`` (the literal expansion of
``pycollo/backend.py:1433-1679``) but organised by blocks so it runs at any mesh
size; it is also the "port" CPU baseline timed by ``bench.py``.

Per phase, with ``Y = V*y_tilde + r`` etc., ``h = (tF - t0)/2``, ``A`` the
difference CSR (``mesh.sA_matrix``), ``I`` the integration CSR
(``mesh.sI_matrix``) and ``Wq = mesh.W_matrix``:

  defect   c[i,r] = W_d[i] * ( sum_m A[r,m] Y[i,m] + h * sum_m I[r,m] f_i(m) )   backend.py:1572-1603
  path     c[j,m] = W_p[j] * p_j(m)                                             backend.py:1605-1616
  integral c[i]   = W_i[i] * ( Q_i - h * sum_m Wq[m] g_i(m) )                   backend.py:1618-1647
  endpoint c[k]   = W_b[k] * b_k(endpoint variables)                            backend.py:1649-1655
  J = w * J(endpoint variables)                                                 backend.py:1495-1504

G = dc/dx_tilde and H = triu(d2/dx_tilde2 (sigma*J + lam.c)) follow by the chain
rule; the block formulas are the ones the reference's legacy assembly spells
out (``pycollo/compiled.py:213-303, 484-500``, ``pycollo/iteration.py:1021-1103``).
Entries are emitted as COO triplets (duplicates allowed), then merged and
ordered column-major = CasADi CCS (``backend.py:1747-1761``).  The *pattern* is
decided symbolically (derivative expression not identically zero, mesh
coefficient not exactly zero when ``prune``), never from numeric values.
"""
from __future__ import annotations

import numpy as np
import sympy as sym

from .common import lower, ocp_variable_offsets


def _nz(expr):
    return sym.sympify(expr) != 0


class _Fn:
    """Vectorised evaluation of a list of expressions over nodes."""

    def __init__(self, args, exprs):
        self.n = len(exprs)
        self.fn = sym.lambdify(list(args), list(exprs), modules="numpy",
                               cse=True) if exprs else None

    def __call__(self, cols, N):
        if not self.n:
            return np.empty((0, N))
        out = self.fn(*cols)
        return np.vstack([np.broadcast_to(np.asarray(o, dtype=float), (N,))
                          for o in out])


class _FnNumba:
    """Same contract as ``_Fn``, but the expressions are evaluated by ONE
    ``numba.njit(parallel=True)`` kernel that loops over the mesh nodes with
    ``prange`` -- the node-parallel CPU baseline of SURVEY.md section 8(d)(ii)
    (the reference's dead stack planned exactly this: ``@nb.njit`` at
    ``pycollo/numbafy.py:232``).  The scalar body is sympy's own code printer
    output after ``cse``; threads = ``numba.get_num_threads()``."""

    _count = 0

    def __init__(self, args, exprs):
        self.n = len(exprs)
        self.kernel = None
        if not exprs:
            return
        import math
        import numba
        from sympy.printing.pycode import pycode
        args = list(args)
        # user symbol names may shadow builtins or be invalid identifiers: rename
        safe = [sym.Symbol(f"v{i}_") for i in range(len(args))]
        exprs = [sym.sympify(e).xreplace(dict(zip(args, safe))) for e in exprs]
        rep, red = sym.cse(exprs, symbols=sym.numbered_symbols("_t"))
        lines = ["def _kernel(N, out, %s):" % ", ".join(f"c{i}" for i in range(len(args))),
                 "    for m in prange(N):"]
        lines += [f"        v{i}_ = c{i}[m]" for i in range(len(args))]
        lines += [f"        {lhs} = {pycode(rhs, fully_qualified_modules=False)}" for lhs, rhs in rep]
        lines += [f"        out[{k}, m] = {pycode(e, fully_qualified_modules=False)}"
                  for k, e in enumerate(red)]
        src = "\n".join(lines)
        ns = {"prange": numba.prange, "math": math}
        ns.update({k: getattr(math, k) for k in ("sin", "cos", "tan", "exp", "log", "sqrt",
                                                  "asin", "acos", "atan", "atan2", "sinh",
                                                  "cosh", "tanh", "pi", "fabs")})
        exec(src, ns)
        _FnNumba._count += 1
        self.kernel = numba.njit(parallel=True, fastmath=False, cache=False)(ns["_kernel"])

    def __call__(self, cols, N):
        if not self.n:
            return np.empty((0, N))
        out = np.empty((self.n, N))
        cols = [np.ascontiguousarray(np.broadcast_to(np.asarray(c, dtype=float), (N,)))
                for c in cols]
        self.kernel(N, out, *cols)
        return out


class _MergePlan:
    """Sort COO triplets column-major and sum duplicates (pattern fixed once)."""

    def __init__(self, rows, cols, num_rows):
        rows = np.asarray(rows, dtype=np.int64)
        cols = np.asarray(cols, dtype=np.int64)
        key = cols * np.int64(num_rows) + rows
        self.order = np.argsort(key, kind="stable")
        skey = key[self.order]
        first = np.ones(len(skey), dtype=bool)
        first[1:] = skey[1:] != skey[:-1]
        self.starts = np.flatnonzero(first)
        self.rows = rows[self.order][self.starts]
        self.cols = cols[self.order][self.starts]

    def merge(self, vals):
        if len(self.starts) == 0:
            return np.zeros(0)
        return np.add.reduceat(vals[self.order], self.starts)


class BlockwiseNLP:
    def __init__(self, ocp, bounds, meshes, *, scaling_method="bounds", w=1.0,
                 W_ocp=None, prune=True, node_eval="numpy"):
        """``node_eval``: "numpy" (vectorised, one thread) or "numba" (one
        ``prange`` kernel over the nodes, all host cores)."""
        _Fn_ = _Fn if node_eval == "numpy" else _FnNumba
        lp = lower(ocp, bounds, scaling_method)
        self.lp = lp
        self.meshes = meshes
        self.prune = prune
        self.offs, n_ocp = ocp_variable_offsets(lp)
        self.s_ocp_off = n_ocp
        self.V, self.r = lp.V_ocp, lp.r_ocp
        self.w = float(w)
        n_c_ocp = sum(len(ph.f) + len(ph.p) + len(ph.g) for ph in lp.phases) \
            + len(lp.b)
        self.W_ocp = np.ones(n_c_ocp) if W_ocp is None else \
            np.asarray(W_ocp, dtype=float)
        assert self.W_ocp.shape == (n_c_ocp,)

        # ---- layouts (backend.py:1433-1457, 1551-1563; iteration.py:196-314)
        self.x_off, self.c_off, self.Wc_off = [], [], []
        xo = co = wo = 0
        for ph, mesh in zip(lp.phases, meshes):
            N = int(mesh["N"])
            self.x_off.append(xo)
            self.c_off.append(co)
            self.Wc_off.append(wo)
            xo += (len(ph.y) + len(ph.u)) * N + len(ph.q) + len(ph.t)
            co += len(ph.f) * (N - 1) + len(ph.p) * N + len(ph.g)
            wo += len(ph.f) + len(ph.p) + len(ph.g)
        self.s_off = xo
        self.num_x = xo + len(lp.s)
        self.b_off = co
        self.num_c = co + len(lp.b)
        self.Wb_off = wo

        # ---- per-phase symbolic pieces ----
        self.ph = []
        for ph, mesh in zip(lp.phases, meshes):
            v = list(ph.y) + list(ph.u)
            sv = list(lp.s)
            allv = v + sv
            fns = list(ph.f) + list(ph.p) + list(ph.g)   # families d, p, i
            fam = ["d"] * len(ph.f) + ["p"] * len(ph.p) + ["i"] * len(ph.g)
            d1 = [(e, a) for e in range(len(fns)) for a in range(len(allv))
                  if _nz(sym.diff(fns[e], allv[a]))]
            d2 = []
            for e in range(len(fns)):
                for a in range(len(allv)):
                    da = sym.diff(fns[e], allv[a])
                    if not _nz(da):
                        continue
                    for b in range(a, len(allv)):
                        if _nz(sym.diff(da, allv[b])):
                            d2.append((e, a, b))
            rec = dict(
                fns=fns, fam=fam, nv=len(v), d1=d1, d2=d2,
                val=_Fn_(allv, fns),
                jac=_Fn_(allv, [sym.diff(fns[e], allv[a]) for e, a in d1]),
                hes=_Fn_(allv, [sym.diff(fns[e], allv[a], allv[b])
                               for e, a, b in d2]),
                nonzero_fn=[_nz(e) for e in fns],
            )
            I = mesh["sI"].tocoo()
            keep = (I.data != 0.0) if prune else np.ones(len(I.data), bool)
            rec["I_row"] = I.row[keep].astype(np.int64)
            rec["I_col"] = I.col[keep].astype(np.int64)
            rec["I_val"] = I.data[keep].astype(float)
            A = mesh["sA"].tocoo()
            rec["A_row"] = A.row.astype(np.int64)
            rec["A_col"] = A.col.astype(np.int64)
            rec["A_val"] = A.data.astype(float)
            Wq = np.asarray(mesh["W"], dtype=float)
            rec["Wq"] = Wq
            rec["Wq_on"] = (Wq != 0.0) if prune else np.ones(len(Wq), bool)
            N = int(mesh["N"])
            act = np.zeros(N, dtype=bool)
            act[rec["I_col"]] = True
            rec["d_on"] = act
            self.ph.append(rec)

        # ---- point functions J, b over the ordered point variables ----
        pts, pidx, pV = [], [], []
        for ip, (ph, off, mesh) in enumerate(zip(lp.phases, self.offs, meshes)):
            N = int(mesh["N"])
            for i, (a, c) in enumerate(zip(ph.y_t0, ph.y_tF)):
                pts += [a, c]
                pidx += [self.x_off[ip] + i * N, self.x_off[ip] + i * N + N - 1]
                pV += [off["y"] + i] * 2
            base = self.x_off[ip] + (len(ph.y) + len(ph.u)) * N
            for i, q in enumerate(ph.q):
                pts.append(q)
                pidx.append(base + i)
                pV.append(off["q"] + i)
            for i, t in enumerate(ph.t):
                pts.append(t)
                pidx.append(base + len(ph.q) + i)
                pV.append(off["t"] + i)
        for i, s in enumerate(lp.s):
            pts.append(s)
            pidx.append(self.s_off + i)
            pV.append(self.s_ocp_off + i)
        self.pts = pts
        self.pidx = np.array(pidx, dtype=np.int64)
        self.pV = np.array(pV, dtype=np.int64)
        pf = [lp.J] + list(lp.b)
        self.pt_val = sym.lambdify(pts, pf, modules="numpy")
        self.pt_d1 = [(e, a) for e in range(len(pf)) for a in range(len(pts))
                      if _nz(sym.diff(pf[e], pts[a]))]
        self.pt_d1_fn = sym.lambdify(
            pts, [sym.diff(pf[e], pts[a]) for e, a in self.pt_d1],
            modules="numpy")
        self.pt_d2 = []
        for e in range(len(pf)):
            for a in range(len(pts)):
                da = sym.diff(pf[e], pts[a])
                if not _nz(da):
                    continue
                for b in range(a, len(pts)):      # pidx is increasing
                    if _nz(sym.diff(da, pts[b])):
                        self.pt_d2.append((e, a, b))
        self.pt_d2_fn = sym.lambdify(
            pts, [sym.diff(pf[e], pts[a], pts[b]) for e, a, b in self.pt_d2],
            modules="numpy")
        self._G_plan = None
        self._H_plan = None

    # ------------------------------------------------------------ helpers --
    def _unpack(self, x):
        x = np.asarray(x, dtype=float)
        out = []
        for ip, (ph, off, mesh) in enumerate(zip(self.lp.phases, self.offs,
                                                 self.meshes)):
            N = int(mesh["N"])
            ny, nu = len(ph.y), len(ph.u)
            xo = self.x_off[ip]
            blk = x[xo:xo + (ny + nu) * N].reshape(ny + nu, N)
            Vv = self.V[off["y"]:off["y"] + ny + nu][:, None]
            rv = self.r[off["y"]:off["y"] + ny + nu][:, None]
            YU = Vv * blk + rv
            base = xo + (ny + nu) * N
            Q = self.V[off["q"]:off["q"] + len(ph.q)] * x[base:base + len(ph.q)] \
                + self.r[off["q"]:off["q"] + len(ph.q)]
            tt = self.V[off["t"]:off["t"] + len(ph.t)] * \
                x[base + len(ph.q):base + len(ph.q) + len(ph.t)] \
                + self.r[off["t"]:off["t"] + len(ph.t)]
            tvals = dict(zip(ph.t, tt))
            t0 = tvals[ph.t0] if ph.t0 in tvals else float(ph.t0)
            tF = tvals[ph.tF] if ph.tF in tvals else float(ph.tF)
            out.append(dict(YU=YU, Q=Q, t0=t0, tF=tF, h=0.5 * (tF - t0), N=N))
        ns = len(self.lp.s)
        S = self.V[self.s_ocp_off:self.s_ocp_off + ns] * x[self.s_off:] \
            + self.r[self.s_ocp_off:self.s_ocp_off + ns]
        return out, S

    def _cols(self, st, S):
        return [st["YU"][k] for k in range(st["YU"].shape[0])] + \
            [np.full(st["N"], s) for s in S]

    def _points(self, x):
        x = np.asarray(x, dtype=float)
        return self.V[self.pV] * x[self.pidx] + self.r[self.pV]

    # ---------------------------------------------------------- callbacks --
    def J(self, x):
        vals = self.pt_val(*self._points(x))
        return self.w * float(vals[0])

    def g(self, x):
        out = np.zeros(self.num_x)
        d = self.pt_d1_fn(*self._points(x))
        for (e, a), v in zip(self.pt_d1, d):
            if e == 0:
                out[self.pidx[a]] += self.w * self.V[self.pV[a]] * float(v)
        return out

    def dy(self, x):
        sts, S = self._unpack(x)
        out = []
        for ph, rec, st in zip(self.lp.phases, self.ph, sts):
            F = rec["val"](self._cols(st, S), st["N"])
            out.append(F[:len(ph.f)].ravel())
        return np.concatenate(out) if out else np.zeros(0)

    def c(self, x):
        sts, S = self._unpack(x)
        c = np.zeros(self.num_c)
        for ip, (ph, rec, st) in enumerate(zip(self.lp.phases, self.ph, sts)):
            N = st["N"]
            F = rec["val"](self._cols(st, S), N)
            ny, npth, nq = len(ph.f), len(ph.p), len(ph.g)
            co, wo = self.c_off[ip], self.Wc_off[ip]
            for i in range(ny):
                Ay = np.bincount(rec["A_row"],
                                 rec["A_val"] * st["YU"][i][rec["A_col"]],
                                 minlength=N - 1)
                If = np.bincount(rec["I_row"],
                                 rec["I_val"] * F[i][rec["I_col"]],
                                 minlength=N - 1)
                c[co + i * (N - 1):co + (i + 1) * (N - 1)] = \
                    self.W_ocp[wo + i] * (Ay + st["h"] * If)
            po = co + ny * (N - 1)
            for j in range(npth):
                c[po + j * N:po + (j + 1) * N] = \
                    self.W_ocp[wo + ny + j] * F[ny + j]
            io = po + npth * N
            for i in range(nq):
                c[io + i] = self.W_ocp[wo + ny + npth + i] * (
                    st["Q"][i] - st["h"] * np.dot(rec["Wq"], F[ny + npth + i]))
        vals = self.pt_val(*self._points(x))
        for k in range(len(self.lp.b)):
            c[self.b_off + k] = self.W_ocp[self.Wb_off + k] * float(vals[1 + k])
        return c

    # ---------------------------------------------------------------- G ----
    def _G_triplets(self, x, pattern_only=False):
        """COO triplets of G; ``pattern_only`` returns only (rows, cols)."""
        rows, cols, vals = [], [], []

        def emit(r, c_, v):
            r = np.atleast_1d(np.asarray(r, dtype=np.int64))
            c_ = np.atleast_1d(np.asarray(c_, dtype=np.int64))
            r, c_ = np.broadcast_arrays(r, c_)
            rows.append(r.ravel())
            cols.append(c_.ravel())
            if not pattern_only:
                vals.append(np.broadcast_to(np.asarray(v, dtype=float),
                                            r.shape).ravel())

        if pattern_only:
            sts, S = None, None
        else:
            sts, S = self._unpack(x)
        for ip, (ph, rec, off, mesh) in enumerate(zip(
                self.lp.phases, self.ph, self.offs, self.meshes)):
            N = int(mesh["N"])
            ny, nu, npth, nq = len(ph.y), len(ph.u), len(ph.p), len(ph.g)
            nv = ny + nu
            xo, co, wo = self.x_off[ip], self.c_off[ip], self.Wc_off[ip]
            base = xo + nv * N
            tcol = {t: base + nq + k for k, t in enumerate(ph.t)}
            tV = {t: self.V[off["t"] + k] for k, t in enumerate(ph.t)}
            if not pattern_only:
                st = sts[ip]
                colsv = self._cols(st, S)
                F = rec["val"](colsv, N)
                D1 = rec["jac"](colsv, N)
                h = st["h"]
            Ir, Ic, Iv = rec["I_row"], rec["I_col"], rec["I_val"]
            Wq, Wq_on = rec["Wq"], rec["Wq_on"]
            nodes = np.arange(N)
            # difference operator on the state's own column block
            for i in range(ny):
                emit(co + i * (N - 1) + rec["A_row"], xo + i * N + rec["A_col"],
                     None if pattern_only else
                     self.W_ocp[wo + i] * self.V[off["y"] + i] * rec["A_val"])
            # first-derivative entries
            for k, (e, a) in enumerate(rec["d1"]):
                famk = rec["fam"][e]
                Wc = self.W_ocp[wo + e]
                if a < nv:
                    Va = self.V[off["y"] + a]
                    ccol = xo + a * N
                else:
                    Va = self.V[self.s_ocp_off + (a - nv)]
                    ccol = self.s_off + (a - nv)
                if famk == "d":
                    i = e
                    if a < nv:
                        emit(co + i * (N - 1) + Ir, ccol + Ic,
                             None if pattern_only else
                             Wc * h * Va * Iv * D1[k][Ic])
                    else:
                        rsum = None if pattern_only else np.bincount(
                            Ir, Iv * D1[k][Ic], minlength=N - 1)
                        emit(co + i * (N - 1) + np.arange(N - 1), ccol,
                             None if pattern_only else Wc * h * Va * rsum)
                elif famk == "p":
                    j = e - ny
                    r0 = co + ny * (N - 1) + j * N
                    if a < nv:
                        emit(r0 + nodes, ccol + nodes,
                             None if pattern_only else Wc * Va * D1[k])
                    else:
                        emit(r0 + nodes, ccol,
                             None if pattern_only else Wc * Va * D1[k])
                else:
                    i = e - ny - npth
                    r0 = co + ny * (N - 1) + npth * N + i
                    if a < nv:
                        sel = nodes[Wq_on]
                        emit(r0, ccol + sel,
                             None if pattern_only else
                             -Wc * h * Va * Wq[sel] * D1[k][sel])
                    else:
                        emit(r0, ccol,
                             None if pattern_only else
                             -Wc * h * Va * np.dot(Wq, D1[k]))
            # time columns
            for e in range(ny + npth + nq):
                if not rec["nonzero_fn"][e] or rec["fam"][e] == "p":
                    continue
                Wc = self.W_ocp[wo + e]
                for tsym, sign in ((ph.t0, -1.0), (ph.tF, 1.0)):
                    if tsym not in tcol:
                        continue
                    if rec["fam"][e] == "d":
                        v = None if pattern_only else \
                            sign * 0.5 * tV[tsym] * Wc * np.bincount(
                                Ir, Iv * F[e][Ic], minlength=N - 1)
                        emit(co + e * (N - 1) + np.arange(N - 1), tcol[tsym], v)
                    else:
                        i = e - ny - npth
                        v = None if pattern_only else \
                            -sign * 0.5 * tV[tsym] * Wc * np.dot(Wq, F[e])
                        emit(co + ny * (N - 1) + npth * N + i, tcol[tsym], v)
            # integral variable on its own row
            for i in range(nq):
                emit(co + ny * (N - 1) + npth * N + i, base + i,
                     None if pattern_only else
                     self.W_ocp[wo + ny + npth + i] * self.V[off["q"] + i])
        # endpoint constraints
        if not pattern_only:
            d = self.pt_d1_fn(*self._points(x))
        for k, (e, a) in enumerate(self.pt_d1):
            if e == 0:
                continue
            emit(self.b_off + e - 1, self.pidx[a],
                 None if pattern_only else
                 self.W_ocp[self.Wb_off + e - 1] * self.V[self.pV[a]] * float(d[k]))
        rows = np.concatenate(rows) if rows else np.zeros(0, np.int64)
        cols = np.concatenate(cols) if cols else np.zeros(0, np.int64)
        if pattern_only:
            return rows, cols
        return rows, cols, np.concatenate(vals) if vals else np.zeros(0)

    def G_structure(self):
        if self._G_plan is None:
            r, c_ = self._G_triplets(None, pattern_only=True)
            self._G_plan = _MergePlan(r, c_, self.num_c)
        return self._G_plan.rows, self._G_plan.cols

    def G_nonzeros(self, x):
        self.G_structure()
        _, _, v = self._G_triplets(x)
        return self._G_plan.merge(v)

    # ---------------------------------------------------------------- H ----
    def _H_triplets(self, x, sigma, lam, pattern_only=False):
        rows, cols, vals = [], [], []

        def emit(r, c_, v):
            r = np.atleast_1d(np.asarray(r, dtype=np.int64))
            c_ = np.atleast_1d(np.asarray(c_, dtype=np.int64))
            r, c_ = np.broadcast_arrays(r, c_)
            lo, hi = np.minimum(r, c_), np.maximum(r, c_)
            rows.append(lo.ravel())
            cols.append(hi.ravel())
            if not pattern_only:
                vals.append(np.broadcast_to(np.asarray(v, dtype=float),
                                            r.shape).ravel())

        if not pattern_only:
            sts, S = self._unpack(x)
            lam = np.asarray(lam, dtype=float)
        for ip, (ph, rec, off, mesh) in enumerate(zip(
                self.lp.phases, self.ph, self.offs, self.meshes)):
            N = int(mesh["N"])
            ny, nu, npth, nq = len(ph.y), len(ph.u), len(ph.p), len(ph.g)
            nv = ny + nu
            xo, co, wo = self.x_off[ip], self.c_off[ip], self.Wc_off[ip]
            base = xo + nv * N
            tcol = {t: base + nq + k for k, t in enumerate(ph.t)}
            tV = {t: self.V[off["t"] + k] for k, t in enumerate(ph.t)}
            Ir, Ic, Iv = rec["I_row"], rec["I_col"], rec["I_val"]
            Wq = rec["Wq"]
            # activity of each function family at each node
            on = {"d": rec["d_on"], "p": np.ones(N, bool), "i": rec["Wq_on"]}
            nodes = np.arange(N)
            if not pattern_only:
                st = sts[ip]
                colsv = self._cols(st, S)
                D1 = rec["jac"](colsv, N)
                D2 = rec["hes"](colsv, N)
                h = st["h"]
                mult = np.zeros((ny + npth + nq, N))   # multiplies F_e(m) in L
                cf = np.ones(ny + npth + nq)           # 1 if scaled by h else 0
                for i in range(ny):
                    lam_i = lam[co + i * (N - 1):co + (i + 1) * (N - 1)]
                    mult[i] = self.W_ocp[wo + i] * np.bincount(
                        Ic, Iv * lam_i[Ir], minlength=N)
                for j in range(npth):
                    r0 = co + ny * (N - 1) + j * N
                    mult[ny + j] = self.W_ocp[wo + ny + j] * lam[r0:r0 + N]
                    cf[ny + j] = 0.0
                for i in range(nq):
                    r0 = co + ny * (N - 1) + npth * N + i
                    mult[ny + npth + i] = \
                        -self.W_ocp[wo + ny + npth + i] * lam[r0] * Wq
                hfac = np.where(cf == 1.0, h, 1.0)

            def Vof(a, off=off, nv=nv):
                return self.V[off["y"] + a] if a < nv else \
                    self.V[self.s_ocp_off + (a - nv)]

            def colof(a, xo=xo, N=N, nv=nv):
                return (xo + a * N + nodes) if a < nv else \
                    np.full(N, self.s_off + (a - nv))

            # second derivatives of phase functions
            for k, (e, a, b) in enumerate(rec["d2"]):
                sel = on[rec["fam"][e]]
                if a < nv and b < nv:
                    emit(colof(a)[sel], colof(b)[sel],
                         None if pattern_only else
                         (Vof(a) * Vof(b) * hfac[e] * mult[e] * D2[k])[sel])
                elif a < nv <= b:
                    emit(colof(a)[sel], colof(b)[sel],
                         None if pattern_only else
                         (Vof(a) * Vof(b) * hfac[e] * mult[e] * D2[k])[sel])
                else:
                    if not sel.any():
                        continue
                    emit(self.s_off + (a - nv), self.s_off + (b - nv),
                         None if pattern_only else
                         Vof(a) * Vof(b) * hfac[e] * np.sum((mult[e] * D2[k])[sel]))
            # mixed derivatives with the free times (through h only)
            for k, (e, a) in enumerate(rec["d1"]):
                famk = rec["fam"][e]
                if famk == "p":
                    continue
                sel = on[famk]
                for tsym, sign in ((ph.t0, -1.0), (ph.tF, 1.0)):
                    if tsym not in tcol:
                        continue
                    if a < nv:
                        emit(colof(a)[sel], np.full(N, tcol[tsym])[sel],
                             None if pattern_only else
                             (sign * 0.5 * tV[tsym] * Vof(a) * mult[e] * D1[k])[sel])
                    else:
                        if not sel.any():
                            continue
                        emit(tcol[tsym], self.s_off + (a - nv),
                             None if pattern_only else
                             sign * 0.5 * tV[tsym] * Vof(a)
                             * np.sum((mult[e] * D1[k])[sel]))
        # endpoint block: sigma*w*J + lam_b.W_b.b
        if not pattern_only:
            d2 = self.pt_d2_fn(*self._points(x))
        for k, (e, a, b) in enumerate(self.pt_d2):
            if pattern_only:
                emit(self.pidx[a], self.pidx[b], None)
                continue
            m = sigma * self.w if e == 0 else \
                lam[self.b_off + e - 1] * self.W_ocp[self.Wb_off + e - 1]
            emit(self.pidx[a], self.pidx[b],
                 m * self.V[self.pV[a]] * self.V[self.pV[b]] * float(d2[k]))
        rows = np.concatenate(rows) if rows else np.zeros(0, np.int64)
        cols = np.concatenate(cols) if cols else np.zeros(0, np.int64)
        if pattern_only:
            return rows, cols
        return rows, cols, np.concatenate(vals) if vals else np.zeros(0)

    def H_structure(self):
        if self._H_plan is None:
            r, c_ = self._H_triplets(None, None, None, pattern_only=True)
            self._H_plan = _MergePlan(r, c_, self.num_x)
        return self._H_plan.rows, self._H_plan.cols

    def H_nonzeros(self, x, sigma, lam):
        self.H_structure()
        _, _, v = self._H_triplets(x, float(sigma), lam)
        return self._H_plan.merge(v)

"""Symbolic derivative analysis of the per-node and point functions.

Decides, once per problem, WHICH derivatives exist (structural sparsity of the
node-local Jacobian/Hessian blocks) and provides their expressions for the CUDA
code generator.  A derivative is structural iff its sympy expression is not
identically zero -- the same criterion the oracle uses (``oracle/expand.py``)
and the closest available stand-in for CasADi's dependency propagation
("bit-exact sparsity vs CasADi without CasADi").

Function families per phase (``pycollo/backend.py:1173-1176`` ordering
``y_eqn + p_con + q_fnc``):  'd' state equations (defect rows), 'p' path
constraints, 'i' integrands (integral rows).
Variables seen by a phase function: ``v = y + u`` (node-local) then ``s``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import sympy as sym


def _nz(e):
    return sym.sympify(e) != 0


@dataclass
class PhaseDerivs:
    index: int
    NY: int
    NU: int
    NP: int
    NQ: int
    NS: int
    variables: list          # y + u + s symbols
    fns: list                # f + p + g expressions
    fam: list                # 'd' / 'p' / 'i' per function
    fn_nonzero: list
    d1v: list = field(default_factory=list)    # (e, a) a < NV
    d1s: list = field(default_factory=list)    # (e, j) j = s index
    d1v_expr: list = field(default_factory=list)
    d1s_expr: list = field(default_factory=list)
    d2: list = field(default_factory=list)     # (e, a, b, expr) a <= b over v+s
    h2vv: list = field(default_factory=list)   # (a, b) a <= b < NV   union over e
    h2vs: list = field(default_factory=list)   # (a, j)
    h2ss: list = field(default_factory=list)   # (i, j) i <= j
    htv: list = field(default_factory=list)    # a with some d/i function dep.
    hts: list = field(default_factory=list)    # j with some d/i function dep.

    NV = property(lambda self: self.NY + self.NU)
    NF = property(lambda self: len(self.fns))

    def pair_families(self, a, b):
        """Families of the functions contributing to second derivative (a, b)."""
        return {self.fam[e] for e, aa, bb, _ in self.d2 if (aa, bb) == (a, b)}

    def d1_families(self, a):
        return {self.fam[e] for e, aa in
                (self.d1v if a < self.NV else
                 [(e, j + self.NV) for e, j in self.d1s]) if aa == a}


# The analysis depends on the symbolic problem only, never on the mesh: a mesh refinement
# (a new Iteration, a new ph-mesh error engine: two lowerings per mesh iteration) finds it
# here instead of differentiating everything again (Delta III: ~1 s per phase, the cse of
# the code generator another 1.5 s).  Keys are the (immutable, hashable) sympy expressions
# themselves; results are read-only by convention.
_CACHE_MAX = 64
_PHASE_CACHE: dict = {}
_POINT_CACHE: dict = {}


def _remember(cache, key, make):
    hit = cache.get(key)
    if hit is None:
        if len(cache) >= _CACHE_MAX:
            cache.pop(next(iter(cache)))
        hit = cache[key] = make()
    return hit


def analyse_phase(ph, s_syms, second=True) -> PhaseDerivs:
    """``second=False`` (``Settings.derivative_level == 1``): no second derivatives
    are formed, so no Hessian program is generated or compiled."""
    key = (ph.index, tuple(ph.y), tuple(ph.u), tuple(sym.sympify(e) for e in ph.f),
           tuple(sym.sympify(e) for e in ph.p), tuple(sym.sympify(e) for e in ph.g),
           tuple(s_syms), bool(second))
    return _remember(_PHASE_CACHE, key, lambda: _analyse_phase(ph, s_syms, second))


def _analyse_phase(ph, s_syms, second) -> PhaseDerivs:
    v = list(ph.y) + list(ph.u)
    allv = v + list(s_syms)
    fns = list(ph.f) + list(ph.p) + list(ph.g)
    fam = ["d"] * len(ph.f) + ["p"] * len(ph.p) + ["i"] * len(ph.g)
    pd = PhaseDerivs(index=ph.index, NY=len(ph.y), NU=len(ph.u), NP=len(ph.p),
                     NQ=len(ph.g), NS=len(s_syms), variables=allv, fns=fns,
                     fam=fam, fn_nonzero=[_nz(e) for e in fns])
    NV = pd.NV
    vv, vs, ss = set(), set(), set()
    for e, fe in enumerate(fns):
        free = fe.free_symbols
        for a, va in enumerate(allv):
            if va not in free:
                continue
            da = sym.diff(fe, va)
            if not _nz(da):
                continue
            if a < NV:
                pd.d1v.append((e, a))
                pd.d1v_expr.append(da)
            else:
                pd.d1s.append((e, a - NV))
                pd.d1s_expr.append(da)
            for b in range(a, len(allv) if second else a):
                if allv[b] not in da.free_symbols:
                    continue
                dab = sym.diff(da, allv[b])
                if not _nz(dab):
                    continue
                pd.d2.append((e, a, b, dab))
                if b < NV:
                    vv.add((a, b))
                elif a < NV:
                    vs.add((a, b - NV))
                else:
                    ss.add((a - NV, b - NV))
    pd.h2vv = sorted(vv, key=lambda t: (t[1], t[0]))   # column-major: by b then a
    pd.h2vs = sorted(vs, key=lambda t: (t[1], t[0]))
    pd.h2ss = sorted(ss, key=lambda t: (t[1], t[0]))
    if second:
        pd.htv = sorted({a for e, a in pd.d1v if fam[e] in "di"})
        pd.hts = sorted({j for e, j in pd.d1s if fam[e] in "di"})
    return pd


@dataclass
class PointDerivs:
    pts: list                 # ordered point symbols
    fns: list                 # [J] + b
    d1: list                  # (e, a)
    d1_expr: list
    d2: list                  # (e, a, b, expr) a <= b
    pairs: list               # (a, b) union over e, a <= b


def analyse_point(ir, second=True) -> PointDerivs:
    key = (tuple(ir.point_symbols), sym.sympify(ir.J), tuple(sym.sympify(e) for e in ir.b),
           bool(second))
    return _remember(_POINT_CACHE, key, lambda: _analyse_point(ir, second))


def _analyse_point(ir, second) -> PointDerivs:
    pts = list(ir.point_symbols)
    fns = [ir.J] + list(ir.b)
    d1, d1e, d2, pairs = [], [], [], set()
    for e, fe in enumerate(fns):
        for a, pa in enumerate(pts):
            if pa not in fe.free_symbols:
                continue
            da = sym.diff(fe, pa)
            if not _nz(da):
                continue
            d1.append((e, a))
            d1e.append(da)
            for b in range(a, len(pts) if second else a):
                if pts[b] not in da.free_symbols:
                    continue
                dab = sym.diff(da, pts[b])
                if _nz(dab):
                    d2.append((e, a, b, dab))
                    pairs.add((a, b))
    return PointDerivs(pts=pts, fns=fns, d1=d1, d1_expr=d1e, d2=d2,
                       pairs=sorted(pairs, key=lambda t: (t[1], t[0])))

"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by ``pycollo_b200``.

Shared lowering of a user problem for the two CPU restatements of the
reference's NLP callbacks (``oracle/expand.py`` -- literal symbolic expansion,
small meshes; ``oracle/blockwise.py`` -- vectorised numpy, any mesh).

The live reference path is CasADi (``pycollo/backend.py:1341-1840``), which
cannot be imported here (casadi, pyproprop absent); the
arithmetic lives in CasADi's SX VM, not under ``/root/reference``.  These files
restate the *published algebra* of that path and are pinned against every
known answer the reference's tests hold for it (``tests/test_oracle_pins.py``):
J, g, c, V, r, x<->x_tilde of ``tests/unit/test_iteration.py:290-385`` and
``tests/unit/test_iteration_scaling.py``.  G and H (no known answers anywhere in the
reference's tests) are pinned against golden vectors produced by
EXECUTING the unmodified reference package under the stand-ins of ``oracle/refshim``
(``oracle/make_golden_nlp.py`` -> ``tests/golden/nlp_*.npz``,
``tests/test_reference_goldens.py``); the two restatements also check each other and are
checked against central differences.

What this module restates
-------------------------
* aux-data resolution: user symbols are replaced through the auxiliary data,
  phase-level entries shadowing problem-level ones, until only root variables
  remain (``backend.py:303-609, 1098-1123``).
* constant folding of variables whose bounds coincide (``bounds.py:456-480``).
* the OCP-level scaling ``V = upper - lower``, ``r = upper - V/2``
  (``scaling.py:87-92``) or ``V = 1, r = 0`` for ``scaling_method`` none
  (``scaling.py:97-101``).
* the variable ordering per phase ``[y.., u.., q.., t..]`` then ``s``
  (``backend.py:1006-1031, 649-662``).

Bound *parsing* (user input formats) is API-layer work outside the hot path
(SURVEY.md §2 row 13); the parsed numeric bounds are taken from the caller.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import sympy as sym


@dataclass
class LoweredPhase:
    y: list
    u: list
    q: list
    t: list            # free time symbols
    t0: object
    tF: object
    y_t0: list
    y_tF: list
    f: list
    p: list
    g: list


@dataclass
class LoweredProblem:
    phases: list
    s: list
    J: object
    b: list
    V_ocp: np.ndarray
    r_ocp: np.ndarray


def _resolve(expr, aux, roots, depth=100):
    expr = sym.sympify(expr)
    for _ in range(depth):
        hit = [s for s in expr.free_symbols if s in aux and s not in roots]
        if not hit:
            return expr
        expr = expr.subs({s: aux[s] for s in hit}, simultaneous=True)
    raise ValueError("recursive auxiliary data")


def lower(ocp, bounds, scaling_method="bounds"):
    """``bounds``: list per phase of dict(y=(n,2), u=, q=, t=(2,2)) + ``bounds['s']``.

    Returns the lowered symbolic problem and the OCP-level V, r vectors.
    """
    problem_aux = {k: sym.sympify(v) for k, v in ocp.auxiliary_data.items()}
    s_user = list(ocp.parameter_variables)
    s_b = np.asarray(bounds["s"], dtype=float).reshape(len(s_user), 2)
    s_keep = [not np.isclose(lo, hi) for lo, hi in s_b]
    consts = {s: sym.Float(0.5 * (lo + hi))
              for s, (lo, hi), k in zip(s_user, s_b, s_keep) if not k}
    s = [v for v, k in zip(s_user, s_keep) if k]
    x_bnd = []
    phases = []
    point_roots = set(s)
    for ph, pb in zip(ocp.phases, bounds["phases"]):
        aux = dict(problem_aux)
        aux.update({k: sym.sympify(v) for k, v in ph.auxiliary_data.items()})
        groups = {}
        pconsts = dict(consts)
        for key, user in (("y", ph.state_variables), ("u", ph.control_variables),
                          ("q", ph.integral_variables), ("t", ph.time_variables)):
            user = list(user)
            bb = np.asarray(pb[key], dtype=float).reshape(len(user), 2)
            keep = [not np.isclose(lo, hi) for lo, hi in bb]
            groups[key] = ([v for v, k in zip(user, keep) if k],
                           bb[np.array(keep, dtype=bool)] if len(user) else bb,
                           keep)
            for v, (lo, hi), k in zip(user, bb, keep):
                if not k:
                    pconsts[v] = sym.Float(0.5 * (lo + hi))
        ykeep = groups["y"][2]
        for v, (lo, hi), k in zip(ph.initial_state_variables,
                                  np.asarray(pb["y"], dtype=float).reshape(-1, 2),
                                  ykeep):
            if not k:
                pconsts[v] = sym.Float(0.5 * (lo + hi))
        for v, (lo, hi), k in zip(ph.final_state_variables,
                                  np.asarray(pb["y"], dtype=float).reshape(-1, 2),
                                  ykeep):
            if not k:
                pconsts[v] = sym.Float(0.5 * (lo + hi))
        consts.update({k: v for k, v in pconsts.items() if k not in consts})
        y, u, q, t = (groups[k][0] for k in "yuqt")
        roots = set(y) | set(u) | set(s)

        def low(e, aux=aux, roots=roots, pconsts=pconsts):
            e = _resolve(e, aux, roots)
            e = e.subs(pconsts, simultaneous=True)
            return _resolve(e, aux, roots).subs(pconsts, simultaneous=True)

        f = [low(e) for e, k in zip(ph.state_equations, ykeep) if k]
        p = [low(e) for e in ph.path_constraints]
        g = [low(e) for e, k in zip(ph.integrand_functions, groups["q"][2]) if k]
        tb = np.asarray(pb["t"], dtype=float).reshape(2, 2)
        tkeep = groups["t"][2]
        t_user = list(ph.time_variables)
        t0 = t_user[0] if tkeep[0] else sym.Float(0.5 * (tb[0, 0] + tb[0, 1]))
        tF = t_user[1] if tkeep[1] else sym.Float(0.5 * (tb[1, 0] + tb[1, 1]))
        y_t0 = [v for v, k in zip(ph.initial_state_variables, ykeep) if k]
        y_tF = [v for v, k in zip(ph.final_state_variables, ykeep) if k]
        point_roots |= set(y_t0) | set(y_tF) | set(q) | set(t)
        phases.append(LoweredPhase(y, u, q, t, t0, tF, y_t0, y_tF, f, p, g))
        for key in "yuqt":
            x_bnd.append(groups[key][1].reshape(-1, 2))
    x_bnd.append(s_b[np.array(s_keep, dtype=bool)].reshape(-1, 2)
                 if len(s_user) else np.empty((0, 2)))
    x_bnd = np.vstack(x_bnd)

    def lowp(e):
        e = _resolve(e, problem_aux, point_roots).subs(consts, simultaneous=True)
        return _resolve(e, problem_aux, point_roots).subs(consts, simultaneous=True)

    J = lowp(ocp.objective_function)
    b = [lowp(e) for e in ocp.endpoint_constraints]
    if scaling_method in (None, "none"):
        V = np.ones(len(x_bnd))
        r = np.zeros(len(x_bnd))
    else:
        lo, hi = x_bnd[:, 0], x_bnd[:, 1]
        V = hi - lo
        r = hi - (hi - lo) / 2
    return LoweredProblem(phases, s, J, b, V, r)


def ocp_variable_offsets(lp):
    """Start index of each phase's [y,u,q,t] block in the OCP-level ordering."""
    offs = []
    k = 0
    for ph in lp.phases:
        d = {"y": k}
        k += len(ph.y)
        d["u"] = k
        k += len(ph.u)
        d["q"] = k
        k += len(ph.q)
        d["t"] = k
        k += len(ph.t)
        offs.append(d)
    return offs, k

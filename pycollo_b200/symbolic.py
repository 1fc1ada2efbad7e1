"""sympy-native preprocessing of the user problem into a flat symbolic IR.

Replaces the CasADi-coupled preprocessing of the reference
(``pycollo/backend.py:149-851`` ``BackendABC`` and ``:854-1305``
``PycolloPhaseData``): user symbols are resolved through the auxiliary data
(phase-level definitions shadow problem-level ones, ``backend.py:1098-1123``)
until only root variables remain, constant variables (equal bounds) are folded
to numbers (``bounds.py:456-480``, ``backend.py:1212-1273``), and the result is
kept in the *unscaled* user basis.  The scaled ("tilde") basis of the NLP,
``x = V * x_tilde + r`` (``backend.py:170-187``), is applied numerically by the
engine: ``d/dx_tilde = V * d/dx``.

The IR is what the CUDA code generator (``codegen.py``) and the sparsity
builder (``structure.py``) consume.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import sympy as sym

_MAX_SUBSTITUTION_DEPTH = 100


@dataclass
class PhaseIR:
    index: int
    name: str
    y: tuple          # unscaled state symbols kept in the NLP
    u: tuple
    q: tuple
    t: tuple          # free time symbols only (subset of (t0, tF))
    t0: object        # Symbol if free else Float
    tF: object
    y_t0: tuple       # endpoint symbols of the kept states
    y_tF: tuple
    f: tuple          # state equations, one per kept state
    p: tuple          # path constraints
    g: tuple          # integrands, one per integral variable
    y_bnd: np.ndarray
    u_bnd: np.ndarray
    q_bnd: np.ndarray
    t_bnd: np.ndarray
    y_t0_bnd: np.ndarray
    y_tF_bnd: np.ndarray
    p_bnd: np.ndarray
    t_needed: tuple
    y_needed: np.ndarray = None
    u_needed: np.ndarray = None
    q_needed: np.ndarray = None
    lower: object = None      # expr -> expr over (y, u, s): aux data and constants resolved

    n_y = property(lambda self: len(self.y))
    n_u = property(lambda self: len(self.u))
    n_q = property(lambda self: len(self.q))
    n_t = property(lambda self: len(self.t))
    n_p = property(lambda self: len(self.p))


@dataclass
class ProblemIR:
    phases: list
    s: tuple
    J: object
    b: tuple
    s_bnd: np.ndarray
    b_bnd: np.ndarray
    s_needed: np.ndarray = None
    point_symbols: tuple = field(default_factory=tuple)
    full_bounds: dict = field(default_factory=dict)   # before constant removal
    lower_point: object = None   # expr -> expr over the endpoint / q / t / s symbols

    n_s = property(lambda self: len(self.s))
    n_b = property(lambda self: len(self.b))


# ---------------------------------------------------------------- bounds ----
def _resolve_number(value, resolver, inf):
    if value is None:
        return None
    if isinstance(value, str):
        if value == "inf":
            return inf
        if value == "-inf":
            return -inf
        return float(value)
    if isinstance(value, (int, float, np.floating, np.integer)):
        return float(value)
    expr = resolver(sym.sympify(value))
    if expr.free_symbols:
        raise ValueError(f"The user-supplied bound '{value}' cannot be "
                         f"precomputed.")
    return float(expr)


def _pair(bnd, resolver, inf):
    flat = np.array(bnd, dtype=object).flatten()
    if flat.shape == (1,):
        lo = hi = _resolve_number(flat[0], resolver, inf)
    elif flat.shape == (2,):
        lo = _resolve_number(flat[0], resolver, inf)
        hi = _resolve_number(flat[1], resolver, inf)
    else:
        raise ValueError(f"Bounds must be single values or (lower, upper) "
                         f"pairs, got '{bnd}'.")
    lo = -inf if lo is None else lo
    hi = inf if hi is None else hi
    return [lo, hi]


def parse_bounds(user_bnds, user_syms, kind, num_expect, settings, resolver,
                 allow_constants=True):
    """(num_expect, 2) array + ``needed`` mask (``bounds.py:560-880``)."""
    inf = settings.numerical_inf
    if num_expect == 0:
        return np.empty((0, 2)), np.empty(0, dtype=bool)
    if user_bnds is None:
        # settings.assume_inf_bounds (reference default): missing -> +-inf
        rows = [[-inf, inf] for _ in range(num_expect)]
    elif isinstance(user_bnds, dict):
        by_name = {str(k): v for k, v in user_bnds.items()}
        rows = []
        for s in user_syms:
            if by_name.get(str(s)) is None:
                rows.append([-inf, inf])
            else:
                rows.append(_pair(by_name[str(s)], resolver, inf))
    else:
        arr = np.array(user_bnds, dtype=object)
        if num_expect == 1 and arr.ndim <= 1 and arr.size in (1, 2):
            rows = [_pair(arr, resolver, inf)]
        else:
            seq = list(user_bnds)
            if len(seq) != num_expect:
                raise ValueError(
                    f"{num_expect} {kind} bounds expected but {len(seq)} "
                    f"supplied.")
            rows = [_pair(b, resolver, inf) for b in seq]
    bnds = np.array(rows, dtype=np.float64)
    same = np.isclose(bnds[:, 0], bnds[:, 1],
                      rtol=settings.bound_clash_absolute_tolerance,
                      atol=settings.bound_clash_relative_tolerance)
    mean = 0.5 * (bnds[:, 0] + bnds[:, 1])
    bnds[same, 0] = mean[same]
    bnds[same, 1] = mean[same]
    bad = bnds[:, 0] > bnds[:, 1]
    if np.any(bad):
        i = int(np.flatnonzero(bad)[0])
        raise ValueError(
            f"The user-supplied upper bound for the {kind} "
            f"'{user_syms[i] if user_syms else i}' (index #{i}) of "
            f"'{bnds[i, 1]}' cannot be less than the user-supplied lower "
            f"bound of '{bnds[i, 0]}'.")
    if allow_constants and settings.remove_constant_variables:
        needed = ~same
    else:
        needed = np.ones(num_expect, dtype=bool)
    return bnds, needed


# ------------------------------------------------------------ resolution ----
class _Resolver:
    """Iterated auxiliary-data substitution for one scope (problem or phase)."""

    def __init__(self, aux, roots):
        self.aux = {k: sym.sympify(v) for k, v in aux.items()}
        self.roots = set(roots)

    def __call__(self, expr):
        expr = sym.sympify(expr)
        for _ in range(_MAX_SUBSTITUTION_DEPTH):
            todo = {s: self.aux[s] for s in expr.free_symbols
                    if s in self.aux and s not in self.roots}
            if not todo:
                break
            expr = expr.xreplace(todo)
        else:
            raise ValueError(f"Auxiliary data for '{expr}' is recursive.")
        return expr

    def check_closed(self, expr, what):
        extra = expr.free_symbols - self.roots
        if extra:
            names = ", ".join(sorted(f"'{s}'" for s in map(str, extra)))
            raise ValueError(f"{names} is not defined (needed by {what}).")


def build_ir(ocp) -> ProblemIR:
    settings = ocp.settings
    s_user = tuple(ocp.parameter_variables)
    problem_aux = dict(ocp.auxiliary_data)

    # problem-level bounds on static parameters decide which stay variables
    const_resolver = _Resolver(problem_aux, ())
    s_bnd_full, s_needed = parse_bounds(
        ocp.bounds.parameter_variables, s_user, "parameter variable",
        len(s_user), settings, const_resolver)
    constants = {s: sym.Float(np.mean(b))
                 for s, b, n in zip(s_user, s_bnd_full, s_needed) if not n}
    s_kept = tuple(s for s, n in zip(s_user, s_needed) if n)

    phases = []
    full_bounds = {"s": s_bnd_full.copy(), "phases": []}
    point_roots = set(s_kept)
    for ph in ocp.phases:
        aux = dict(problem_aux)
        aux.update(ph.auxiliary_data)      # phase-level shadows problem-level
        y_user = tuple(ph.state_variables)
        u_user = tuple(ph.control_variables)
        q_user = tuple(ph.integral_variables)
        t_user = tuple(ph.time_variables)
        bres = _Resolver(aux, ())
        b = ph.bounds
        y_bnd, y_need = parse_bounds(b.state_variables, y_user,
                                     "state variable", len(y_user), settings, bres)
        u_bnd, u_need = parse_bounds(b.control_variables, u_user,
                                     "control variable", len(u_user), settings,
                                     bres)
        q_bnd, q_need = parse_bounds(b.integral_variables, q_user,
                                     "integral variable", len(q_user), settings,
                                     bres)
        t_bnd, t_need = parse_bounds([b.initial_time, b.final_time], t_user,
                                     "time variable", 2, settings, bres)
        if t_bnd[0, 0] > t_bnd[1, 0] or t_bnd[0, 1] > t_bnd[1, 1]:
            raise ValueError(
                f"The bounds for the final time must be greater than the "
                f"bounds for the initial time in phase {ph.name} "
                f"(index #{ph.phase_number}).")
        n_p = len(ph.path_constraints)
        p_bnd, _ = parse_bounds(b.path_constraints, [None] * n_p,
                                "path constraints", n_p, settings, bres,
                                allow_constants=False)
        y_t0_bnd, _ = parse_bounds(b.initial_state_constraints, y_user,
                                   "initial state constraint", len(y_user),
                                   settings, bres, allow_constants=False)
        y_tF_bnd, _ = parse_bounds(b.final_state_constraints, y_user,
                                   "final state constraint", len(y_user),
                                   settings, bres, allow_constants=False)
        if settings.override_endpoint_bounds:
            for arr in (y_t0_bnd, y_tF_bnd):
                arr[:, 0] = np.maximum(arr[:, 0], y_bnd[:, 0])
                arr[:, 1] = np.minimum(arr[:, 1], y_bnd[:, 1])

        full_bounds["phases"].append(dict(y=y_bnd.copy(), u=u_bnd.copy(),
                                          q=q_bnd.copy(), t=t_bnd.copy()))
        phase_consts = dict(constants)
        for syms_, bnds_, need_ in ((y_user, y_bnd, y_need),
                                    (u_user, u_bnd, u_need),
                                    (q_user, q_bnd, q_need),
                                    (t_user, t_bnd, t_need)):
            for s_, b_, n_ in zip(syms_, bnds_, need_):
                if not n_:
                    phase_consts[s_] = sym.Float(np.mean(b_))
        for s_, b_, n_ in zip(ph.initial_state_variables, y_bnd, y_need):
            if not n_:
                phase_consts[s_] = sym.Float(np.mean(b_))
        for s_, b_, n_ in zip(ph.final_state_variables, y_bnd, y_need):
            if not n_:
                phase_consts[s_] = sym.Float(np.mean(b_))
        constants.update({k: v for k, v in phase_consts.items()
                          if k in set(q_user) | set(t_user)
                          | set(ph.initial_state_variables)
                          | set(ph.final_state_variables)})

        y = tuple(s_ for s_, n_ in zip(y_user, y_need) if n_)
        u = tuple(s_ for s_, n_ in zip(u_user, u_need) if n_)
        q = tuple(s_ for s_, n_ in zip(q_user, q_need) if n_)
        t = tuple(s_ for s_, n_ in zip(t_user, t_need) if n_)
        roots = set(y) | set(u) | set(s_kept)
        res = _Resolver(aux, roots | set(q) | set(t))

        def lower(expr, what, res=res, roots=roots, consts=phase_consts):
            e = res(expr).xreplace(consts)
            e = res(e).xreplace(consts)
            res.check_closed(e, what)
            bad = e.free_symbols - roots
            if bad:
                raise NotImplementedError(
                    f"{what} depends on {sorted(map(str, bad))}: integral/time "
                    f"variables inside phase functions are not supported.")
            return e

        f = tuple(lower(e, f"state equation #{i} of phase {ph.name}")
                  for i, (e, n_) in enumerate(zip(ph.state_equations, y_need))
                  if n_)
        p = tuple(lower(e, f"path constraint #{i} of phase {ph.name}")
                  for i, e in enumerate(ph.path_constraints))
        g = tuple(lower(e, f"integrand #{i} of phase {ph.name}")
                  for i, (e, n_) in enumerate(zip(ph.integrand_functions,
                                                  q_need)) if n_)
        y_t0 = tuple(s_ for s_, n_ in zip(ph.initial_state_variables, y_need)
                     if n_)
        y_tF = tuple(s_ for s_, n_ in zip(ph.final_state_variables, y_need)
                     if n_)
        t0 = t_user[0] if t_need[0] else sym.Float(np.mean(t_bnd[0]))
        tF = t_user[1] if t_need[1] else sym.Float(np.mean(t_bnd[1]))
        point_roots |= set(y_t0) | set(y_tF) | set(q) | set(t)
        phases.append(PhaseIR(
            index=ph.phase_number, name=ph.name, y=y, u=u, q=q, t=t, t0=t0,
            tF=tF, y_t0=y_t0, y_tF=y_tF, f=f, p=p, g=g,
            y_bnd=y_bnd[y_need], u_bnd=u_bnd[u_need], q_bnd=q_bnd[q_need],
            t_bnd=t_bnd[t_need], y_t0_bnd=y_t0_bnd[y_need],
            y_tF_bnd=y_tF_bnd[y_need], p_bnd=p_bnd,
            t_needed=(bool(t_need[0]), bool(t_need[1])),
            y_needed=y_need, u_needed=u_need, q_needed=q_need, lower=lower))

    # problem-level expressions see every phase's endpoint symbols
    res = _Resolver(problem_aux, point_roots)

    def lower_point(expr, what):
        e = res(expr).xreplace(constants)
        e = res(e).xreplace(constants)
        res.check_closed(e, what)
        return e

    J = lower_point(ocp.objective_function, "the objective function")
    b_exprs = []
    for i, e in enumerate(ocp.endpoint_constraints):
        e = lower_point(e, f"endpoint constraint #{i}")
        if e in point_roots:
            raise ValueError(
                f"Pycollo cannot automatically transform point constraints to "
                f"state endpoint constraints. Use state endpoint constraints "
                f"for '{e}'.")
        b_exprs.append(e)
    n_b = len(b_exprs)
    b_bnd, _ = parse_bounds(ocp.bounds.endpoint_constraints, [None] * n_b,
                            "endpoint constraints", n_b, settings,
                            const_resolver, allow_constants=False)
    ordered_points = []
    for ph in phases:
        for a, c in zip(ph.y_t0, ph.y_tF):
            ordered_points += [a, c]
        ordered_points += list(ph.q) + list(ph.t)
    ordered_points += list(s_kept)
    return ProblemIR(phases=phases, s=s_kept, J=J, b=tuple(b_exprs),
                     s_bnd=s_bnd_full[s_needed], b_bnd=b_bnd, s_needed=s_needed,
                     point_symbols=tuple(ordered_points),
                     full_bounds=full_bounds, lower_point=lower_point)

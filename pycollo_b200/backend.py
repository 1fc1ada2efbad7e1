"""``backend="cuda"``: the drop-in for pycollo's CasADi evaluation path.

Mirrors, for the per-iterate callback path only, the surface the rest of pycollo
calls on its backend:

* ``Cuda`` stands where ``pycollo.backend.Casadi`` stands
  (``pycollo/backend.py:1341-1840``): ``create_bounds / create_scaling /
  create_quadrature / create_initial_mesh / create_guess /
  create_mesh_iterations / new_mesh_iteration`` (``backend.py:625-851``),
  ``generate_nlp_function_callables`` (``:1403-1411``), ``create_nlp_solver``
  (``:1681-1693``) and the per-callback API ``evaluate_J / evaluate_g /
  evaluate_c / evaluate_G / evaluate_G_nonzeros / evaluate_G_structure /
  evaluate_G_num_nonzero`` (``:1713-1771``) plus the four ``evaluate_H*`` the
  reference leaves as ``NotImplementedError`` (``:1773-1805``).
* ``Iteration`` mirrors ``pycollo/iteration.py:18-653`` up to (not including)
  ``solve``: guess interpolation ``:86-194``, counts and slices ``:196-342``,
  scaling ``:344-373``, NLP generation ``:375-394``, bounds ``:396-453``.
* ``IterationScaling`` mirrors ``pycollo/scaling.py:124-454`` with the
  constraint scaling computed from the *sparse* Jacobian (row norms over the
  CCS values) instead of the dense ``np.array(G)`` of ``scaling.py:394``.

All numbers come from the CUDA engine (``engine.py`` -> ``libpcx.so``); nothing
here evaluates the NLP functions on the CPU.
"""
from __future__ import annotations

from timeit import default_timer as timer
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sparse

from . import codegen, engine as _engine
from .derivs import analyse_phase, analyse_point
from .mesh import Mesh, PhaseMeshData
from .quadrature import Quadrature
from .structure import NLPStructure
from .symbolic import build_ir

NlpResult = SimpleNamespace


def lower_problem(ocp, meshes=None, share_bodies=None, **structure_kwargs):
    """Symbolic lowering + structure + generated header for one mesh.  ``share_bodies``:
    phases whose expression bodies differ in literals only share one instantiation of
    the tile function (``codegen.share_groups``; None = the default, on)."""
    ir = build_ir(ocp)
    # Settings.derivative_level == 1: first derivatives only -- no Hessian program is
    # generated, compiled or exposed (the reference's live backend ignores the
    # setting; its cyipopt plumbing honours it, nlp.py:61)
    second = int(getattr(ocp.settings, "derivative_level", 2)) >= 2
    pds = [analyse_phase(ph, ir.s, second) for ph in ir.phases]
    ptd = analyse_point(ir, second)
    if meshes is None:
        quad = Quadrature(ocp.settings.quadrature_method)
        meshes = [PhaseMeshData(quad, ph.mesh, 2, 20) for ph in ocp.phases]
    S = NLPStructure(ir, pds, ptd, meshes,
                     prune=getattr(ocp.settings, "prune_zero_quadrature_coefficients", True),
                     **structure_kwargs)
    header, layouts = codegen.generate(ir, pds, ptd, S, share=share_bodies)
    return SimpleNamespace(ir=ir, pds=pds, ptd=ptd, S=S, header=header,
                           layouts=layouts, meshes=meshes)


class Bounds:
    """Numeric variable bounds in OCP ordering (``pycollo/bounds.py:414-454``)."""

    def __init__(self, ir):
        rows = []
        for ph in ir.phases:
            rows += [ph.y_bnd, ph.u_bnd, ph.q_bnd, ph.t_bnd]
        rows.append(ir.s_bnd)
        self.x_bnd = np.vstack([r.reshape(-1, 2) for r in rows])
        self.y_t0_bnd = [ph.y_t0_bnd for ph in ir.phases]
        self.y_tF_bnd = [ph.y_tF_bnd for ph in ir.phases]

    x_bnd_lower = property(lambda self: self.x_bnd[:, 0])
    x_bnd_upper = property(lambda self: self.x_bnd[:, 1])


class Scaling:
    """OCP-level stretch/shift (``pycollo/scaling.py:66-115``)."""

    def __init__(self, backend):
        method = backend.ocp.settings.scaling_method
        if method in (None, "none"):
            n = backend.num_var
            self.x_scales, self.x_shifts = np.ones(n), np.zeros(n)
        else:
            lo, hi = backend.bounds.x_bnd_lower, backend.bounds.x_bnd_upper
            self.x_scales = hi - lo
            self.x_shifts = hi - (hi - lo) / 2


class Guess:
    """Initial guess in OCP ordering (``pycollo/guess.py:117-200``)."""

    def __init__(self, backend):
        ir = backend.ir
        self.tau, self.t0, self.tF, self.y, self.u, self.q, self.t = ([] for _ in range(7))
        from .symbolic import _Resolver
        problem_aux = dict(backend.ocp.auxiliary_data)
        for ph, uph in zip(ir.phases, backend.ocp.phases):
            g = uph.guess
            aux = dict(problem_aux)
            aux.update(uph.auxiliary_data)
            self._resolver = _Resolver(aux, ())
            if g.time is None:
                raise ValueError("A guess must be supplied.")
            time = np.asarray(g.time, dtype=np.float64)
            if time.ndim != 1:
                raise ValueError("Time guess must be a 1d array.")
            t0, tF = time[0], time[-1]
            stretch, shift = 0.5 * (tF - t0), 0.5 * (t0 + tF)
            self.tau.append((time - shift) / stretch)
            self.t0.append(t0)
            self.tF.append(tF)
            y = self._check(g.state_variables, len(uph.state_variables), time.size)
            u = self._check(g.control_variables, len(uph.control_variables), time.size)
            q = self._check(g.integral_variables, len(uph.integral_variables))
            self.y.append(y[ph.y_needed])
            self.u.append(u[ph.u_needed])
            self.q.append(q[ph.q_needed])
            self.t.append(np.array([t0, tF])[np.array(ph.t_needed, dtype=bool)])
        self._resolver = _Resolver(problem_aux, ())
        s = self._check(backend.ocp.guess.parameter_variables,
                        len(backend.ocp.parameter_variables))
        self.s = s[ir.s_needed] if len(s) else s

    def _numeric(self, expr):
        """A symbolic guess entry resolved through the auxiliary data and evaluated
        in double precision operation by operation, as the reference's
        ``expr_as_numeric`` does (``pycollo/backend.py:1382-1384``, CasADi ``DM``) --
        not exactly-then-rounded, which can differ in the last bit."""
        import sympy as sym
        resolved = self._resolver(expr)
        if resolved.free_symbols:
            raise ValueError(f"guess entry '{expr}' does not resolve to a number")
        return float(sym.lambdify([], resolved, modules="math")())

    def _check(self, guess, num_var, num_t=None):
        """Shape check + numeric resolution of symbolic guesses through the
        auxiliary data (``pycollo/guess.py:180-200``)."""
        if guess is None or num_var == 0:
            if num_var != 0:
                raise ValueError("A guess must be supplied.")
            return np.empty((0, num_t)) if num_t is not None else np.empty((0,))
        raw = np.asarray(guess, dtype=object)
        flat = np.array([v if isinstance(v, (int, float, np.floating, np.integer))
                         else self._numeric(v) for v in raw.ravel()], dtype=np.float64)
        guess = flat.reshape(raw.shape)
        if num_t is not None:
            if guess.shape != (num_var, num_t):
                raise ValueError("A guess must be supplied for every symbol and time.")
        else:
            guess = guess.reshape(-1)
            if guess.shape != (num_var,):
                raise ValueError("A guess must be supplied for every symbol.")
        return guess


class IterationScaling:
    """Mesh-expanded V, r and the objective/constraint scaling of one iteration."""

    def __init__(self, iteration):
        self.iteration = iteration
        self.backend = iteration.backend
        base = self.backend.scaling
        self.V_ocp = base.x_scales.copy()
        self.r_ocp = base.x_shifts.copy()
        self.V = self._expand_x_to_mesh(self.V_ocp)
        self.r = self._expand_x_to_mesh(self.r_ocp)
        self.V_inv = np.reciprocal(self.V)
        self.w = 1.0
        self.W_ocp = np.ones(self.backend.num_c)

    def scale_x(self, x):
        return np.multiply(self.V_inv, (x - self.r))

    def unscale_x(self, x_tilde):
        return np.multiply(self.V, x_tilde) + self.r

    def unscale_J(self, J_tilde):
        return (1 / self.w) * J_tilde

    def _expand_x_to_mesh(self, base):
        it, S = self.iteration, self.iteration.S
        out = np.empty(S.num_x)
        for ph, t in zip(self.backend.ir.phases, S.ph):
            o = t.V_off
            nv = ph.n_y + ph.n_u
            out[t.x_off:t.x_off + nv * t.N] = np.repeat(base[o["y"]:o["y"] + nv], t.N)
            out[t.q_col:t.q_col + ph.n_q] = base[o["q"]:o["q"] + ph.n_q]
            out[t.q_col + ph.n_q:t.q_col + ph.n_q + ph.n_t] = \
                base[o["t"]:o["t"] + ph.n_t]
        out[S.s_off:] = base[S.s_ocp_off:S.s_ocp_off + S.NS]
        return out

    def _expand_c_to_mesh(self, base):
        S = self.iteration.S
        out = np.empty(S.num_c)
        for ph, t in zip(self.backend.ir.phases, S.ph):
            nd = ph.n_y * (t.N - 1)
            out[t.c_off:t.c_off + nd] = np.repeat(base[t.W_off:t.W_off + ph.n_y], t.N - 1)
            out[t.c_off + nd:t.c_off + nd + ph.n_p * t.N] = \
                np.repeat(base[t.W_off + ph.n_y:t.W_off + ph.n_y + ph.n_p], t.N)
            out[t.c_off + nd + ph.n_p * t.N:t.c_off + nd + ph.n_p * t.N + ph.n_q] = \
                base[t.W_off + ph.n_y + ph.n_p:t.W_off + ph.n_y + ph.n_p + ph.n_q]
        out[S.b_off:] = base[S.Wb_off:S.Wb_off + S.NB]
        return out

    @property
    def W(self):
        return self._expand_c_to_mesh(self.W_ocp)

    def generate_J_c_scaling(self):
        """``pycollo/scaling.py:204-210, 261-275, 346-430`` with sparse row norms."""
        it = self.iteration
        method = self.backend.ocp.settings.scaling_method
        if method in (None, "none"):
            self.w, self.W_ocp = 1.0, np.ones(self.backend.num_c)
            it.push_scaling()
            return
        # evaluate g and G at the guess with unit scaling (w = 1, W = 1)
        self.w, self.W_ocp = 1.0, np.ones(self.backend.num_c)
        it.push_scaling()
        x0 = it.guess_x_tilde
        if it.number == 1:
            w = 1.0
        else:
            g = it.evaluate(_engine.EVAL_GRAD, x0)["grad"][0]
            g_norm = np.sqrt(np.sum(g ** 2))
            w = 1.0 if np.isclose(g_norm, 0.0) else 1.0 / g_norm
        # row norms reduced on the device (pcx_jac_row_norms): num_c doubles come
        # back instead of the nnz_G Jacobian values
        G_norm = it.create_engine().jac_row_norms_host(x0)
        W = np.empty(self.backend.num_c)
        S = it.S
        for ph, t in zip(self.backend.ir.phases, S.ph):
            o = t.V_off
            W[t.W_off:t.W_off + ph.n_y] = np.reciprocal(self.V_ocp[o["y"]:o["y"] + ph.n_y])
            p0 = t.c_off + ph.n_y * (t.N - 1)
            if ph.n_p:
                W[t.W_off + ph.n_y:t.W_off + ph.n_y + ph.n_p] = np.reciprocal(
                    np.mean(G_norm[p0:p0 + ph.n_p * t.N].reshape(ph.n_p, t.N), axis=1))
            W[t.W_off + ph.n_y + ph.n_p:t.W_off + ph.n_y + ph.n_p + ph.n_q] = \
                np.reciprocal(self.V_ocp[o["q"]:o["q"] + ph.n_q])
        W[S.Wb_off:] = np.reciprocal(G_norm[S.b_off:])
        self.w, self.W_ocp = w, W
        it.push_scaling()


def _interp1d_linear(x, y, x_new):
    """``scipy.interpolate.interp1d(x, y, bounds_error=False,
    fill_value="extrapolate")(x_new)`` as the reference calls it
    (``pycollo/iteration.py:128-134``): left bisection clipped to [1, M-1], then
    scipy's ``_call_linear`` formula ``((t - x_lo)/(x_hi - x_lo)) * y_hi +
    ((x_hi - t)/(x_hi - x_lo)) * y_lo`` with the same operation order (bit-identical
    results).  Same rule as ``pcx_interp_guess``; host mirror for deferred use."""
    x, y, x_new = (np.asarray(a, dtype=np.float64) for a in (x, y, x_new))
    hi = np.clip(np.searchsorted(x, x_new), 1, len(x) - 1)
    lo = hi - 1
    den = x[hi] - x[lo]
    return ((x_new - x[lo]) / den) * y[hi] + ((x[hi] - x_new) / den) * y[lo]


class Iteration:
    """One mesh iteration = one NLP (``pycollo/iteration.py:18-653``)."""

    def __init__(self, backend, index, mesh, guess, batch=1, device=0):
        self.backend = backend
        self.ocp = backend.ocp
        self.index = index
        self.number = index + 1
        self.mesh = mesh
        self.prev_guess = guess
        self.batch = batch
        self.device = device
        self.engine = None
        self.initialise()

    # -- initialise (iteration.py:69-79) ----------------------------------
    def initialise(self):
        """Same stages as the reference: guess onto the mesh, counts/slices,
        variable scaling, scaled guess, NLP generation (engine + J/c scaling),
        bounds.  ``settings.defer_engine = True`` stops before anything touches
        the device (structure-only use on a machine without a GPU: CPU tests,
        table inspection); evaluation then starts with ``generate_nlp()``."""
        t0 = timer()
        low = lower_problem(self.ocp, self.mesh.p)
        self.low, self.S = low, low.S
        self.deferred = bool(getattr(self.ocp.settings, "defer_engine", False))
        self.create_variable_constraint_counts_slices()
        self.scaling = IterationScaling(self)
        self.interpolate_guess_to_mesh(self.prev_guess)
        self.guess_x_tilde = self.scaling.scale_x(self.guess_x)
        if not self.deferred:
            self.generate_nlp()
        self.generate_bounds()
        self._time_initialise = timer() - t0

    def interpolate_guess_to_mesh(self, prev):
        """Linear interpolation of the previous guess onto this mesh
        (``iteration.py:86-194``): on the device (``pcx_interp_guess``, one thread
        per variable and node, interp1d's own formula and rounding); the host
        mirror below is only used by a deferred iteration."""
        self.guess_tau = self.mesh.tau
        if not self.deferred:
            parts = []
            for ip in range(len(self.mesh.tau)):
                parts += [np.ravel(prev.y[ip]), np.ravel(prev.u[ip]),
                          np.ravel(prev.q[ip]), np.ravel(prev.t[ip])]
            parts.append(np.ravel(prev.s))
            x_prev = np.concatenate(parts).astype(np.float64)
            self.guess_x = self.create_engine().interp_guess_host(
                x_prev, [np.asarray(t, dtype=np.float64) for t in prev.tau],
                [np.asarray(t, dtype=np.float64) for t in self.mesh.tau])
            self.guess_y = [self.guess_x[sl].reshape(-1, len(tau))
                            for sl, tau in zip(self.y_slices, self.mesh.tau)]
            self.guess_u = [self.guess_x[sl].reshape(-1, len(tau))
                            for sl, tau in zip(self.u_slices, self.mesh.tau)]
            return
        parts = []
        self.guess_y, self.guess_u = [], []
        for ip, (tau, ptau) in enumerate(zip(self.mesh.tau, prev.tau)):
            y = np.vstack([_interp1d_linear(ptau, row, tau) for row in prev.y[ip]]) \
                if len(prev.y[ip]) else np.empty((0, len(tau)))
            u = np.vstack([_interp1d_linear(ptau, row, tau) for row in prev.u[ip]]) \
                if len(prev.u[ip]) else np.empty((0, len(tau)))
            self.guess_y.append(y)
            self.guess_u.append(u)
            parts += [y.ravel(), u.ravel(), np.ravel(prev.q[ip]), np.ravel(prev.t[ip])]
        parts.append(np.ravel(prev.s))
        self.guess_x = np.concatenate(parts).astype(np.float64)

    def create_variable_constraint_counts_slices(self):
        """``iteration.py:196-342``."""
        S, ir = self.S, self.backend.ir
        self.num_x, self.num_c = S.num_x, S.num_c
        self.y_slices, self.u_slices, self.q_slices, self.t_slices = [], [], [], []
        self.x_slices, self.c_defect_slices, self.c_path_slices = [], [], []
        self.c_integral_slices, self.c_slices, self.dy_slices = [], [], []
        for ph, t in zip(ir.phases, S.ph):
            y0 = t.x_off
            u0 = y0 + ph.n_y * t.N
            self.y_slices.append(slice(y0, u0))
            self.u_slices.append(slice(u0, t.q_col))
            self.q_slices.append(slice(t.q_col, t.q_col + ph.n_q))
            self.t_slices.append(slice(t.q_col + ph.n_q, t.q_col + ph.n_q + ph.n_t))
            self.x_slices.append(slice(y0, t.q_col + ph.n_q + ph.n_t))
            d1 = t.c_off + ph.n_y * (t.N - 1)
            p1 = d1 + ph.n_p * t.N
            self.c_defect_slices.append(slice(t.c_off, d1))
            self.c_path_slices.append(slice(d1, p1))
            self.c_integral_slices.append(slice(p1, p1 + ph.n_q))
            self.c_slices.append(slice(t.c_off, p1 + ph.n_q))
            self.dy_slices.append(slice(t.dy_off, t.dy_off + ph.n_y * t.N))
        self.s_slice = slice(S.s_off, S.num_x)
        self.c_endpoint_slice = slice(S.b_off, S.num_c)
        self.num_s = S.NS

    def generate_bounds(self):
        """x and c bounds on the mesh (``iteration.py:396-453``): state endpoint
        constraints are *variable bounds*, not rows of c (``:419-420``)."""
        S, ir = self.S, self.backend.ir
        sc = self.scaling
        if not self.deferred:
            # on the device (pcx_expand_bounds): OCP-level (lo, hi) pairs in, scaled
            # mesh vectors out; c bounds are produced with W = 1 and scaled by the
            # current W in the c_bnd_l / c_bnd_u properties
            c_rows = []
            for ph in ir.phases:
                c_rows += [np.zeros((ph.n_y, 2)), ph.p_bnd.reshape(-1, 2), np.zeros((ph.n_q, 2))]
            c_rows.append(ir.b_bnd.reshape(-1, 2))
            y0 = np.vstack([ph.y_t0_bnd.reshape(-1, 2) for ph in ir.phases])
            yF = np.vstack([ph.y_tF_bnd.reshape(-1, 2) for ph in ir.phases])
            self.x_bnd_l, self.x_bnd_u, cl, cu = self.create_engine().expand_bounds_host(
                self.backend.bounds.x_bnd, y0, yF, np.vstack(c_rows), sc.V_ocp, sc.r_ocp,
                np.ones(self.backend.num_c))
            self._c_bnd_unscaled = (cl, cu)
            return
        lo = np.empty(S.num_x)
        hi = np.empty(S.num_x)
        for ph, t in zip(ir.phases, S.ph):
            for i in range(ph.n_y):
                sl = slice(t.x_off + i * t.N, t.x_off + (i + 1) * t.N)
                lo[sl], hi[sl] = ph.y_bnd[i]
                lo[sl.start], hi[sl.start] = ph.y_t0_bnd[i]
                lo[sl.stop - 1], hi[sl.stop - 1] = ph.y_tF_bnd[i]
            for j in range(ph.n_u):
                sl = slice(t.x_off + (ph.n_y + j) * t.N, t.x_off + (ph.n_y + j + 1) * t.N)
                lo[sl], hi[sl] = ph.u_bnd[j]
            lo[t.q_col:t.q_col + ph.n_q] = ph.q_bnd[:, 0]
            hi[t.q_col:t.q_col + ph.n_q] = ph.q_bnd[:, 1]
            lo[t.q_col + ph.n_q:t.q_col + ph.n_q + ph.n_t] = ph.t_bnd[:, 0]
            hi[t.q_col + ph.n_q:t.q_col + ph.n_q + ph.n_t] = ph.t_bnd[:, 1]
        lo[S.s_off:] = ir.s_bnd[:, 0]
        hi[S.s_off:] = ir.s_bnd[:, 1]
        self.x_bnd_l = sc.scale_x(lo)
        self.x_bnd_u = sc.scale_x(hi)
        cl = np.zeros(S.num_c)
        cu = np.zeros(S.num_c)
        for ph, t, sl in zip(ir.phases, S.ph, self.c_path_slices):
            if ph.n_p:
                cl[sl] = np.repeat(ph.p_bnd[:, 0], t.N)
                cu[sl] = np.repeat(ph.p_bnd[:, 1], t.N)
        cl[S.b_off:] = ir.b_bnd[:, 0]
        cu[S.b_off:] = ir.b_bnd[:, 1]
        self._c_bnd_unscaled = (cl, cu)

    @property
    def c_bnd_l(self):
        return self.scaling.W * self._c_bnd_unscaled[0]

    @property
    def c_bnd_u(self):
        return self.scaling.W * self._c_bnd_unscaled[1]

    # -- the engine --------------------------------------------------------
    def generate_nlp(self):
        """``iteration.py:375-394``: compile callbacks, then derive J/c scaling."""
        t0 = timer()
        self.backend.generate_nlp_function_callables(self)
        self.scaling.generate_J_c_scaling()
        self.backend.create_nlp_solver()
        self._time_generate_nlp = timer() - t0

    def create_engine(self):
        if self.engine is None:
            self.engine = _engine.Engine(self.S, self.low.layouts, self.low.header,
                                         batch=self.batch, device=self.device)
            self.push_scaling()
        return self.engine

    def push_scaling(self):
        if self.engine is not None:
            self.engine.set_scaling(self.scaling.V_ocp, self.scaling.r_ocp,
                                    self.scaling.W_ocp, self.scaling.w)

    def evaluate(self, what, x, lam=None, sigma=None):
        if (what & _engine.EVAL_HESS) and int(self.ocp.settings.derivative_level) < 2:
            raise ValueError("derivative_level=1: no Hessian callback has been generated "
                             "(set settings.derivative_level = 2 for exact Hessians)")
        return self.create_engine().eval_host(what, x, lam, sigma)


class Cuda:
    """The ``backend="cuda"`` object held as ``ocp._backend``."""

    def __init__(self, ocp):
        self.ocp = ocp
        self.ir = build_ir(ocp)
        self.p = self.ir.phases
        self.num_phases = len(self.p)
        self.num_var = sum(ph.n_y + ph.n_u + ph.n_q + ph.n_t for ph in self.p) + self.ir.n_s
        self.num_c = sum(ph.n_y + ph.n_p + ph.n_q for ph in self.p) + self.ir.n_b
        self.num_s_var = self.ir.n_s
        self.num_b_con = self.ir.n_b
        self.mesh_iterations = []
        self.current_iteration = None
        self._slices()

    def _slices(self):
        """OCP-level slices (``pycollo/backend.py:664-704, 781-816``)."""
        self.phase_y_var_slices, self.phase_u_var_slices = [], []
        self.phase_q_var_slices, self.phase_t_var_slices = [], []
        self.phase_variable_slices = []
        self.phase_y_eqn_slices, self.phase_p_con_slices = [], []
        self.phase_q_fnc_slices, self.phase_c_slices = [], []
        v = c = 0
        for ph in self.p:
            self.phase_y_var_slices.append(slice(v, v + ph.n_y))
            self.phase_u_var_slices.append(slice(v + ph.n_y, v + ph.n_y + ph.n_u))
            q0 = v + ph.n_y + ph.n_u
            self.phase_q_var_slices.append(slice(q0, q0 + ph.n_q))
            self.phase_t_var_slices.append(slice(q0 + ph.n_q, q0 + ph.n_q + ph.n_t))
            self.phase_variable_slices.append(slice(v, q0 + ph.n_q + ph.n_t))
            v = q0 + ph.n_q + ph.n_t
            self.phase_y_eqn_slices.append(slice(c, c + ph.n_y))
            self.phase_p_con_slices.append(slice(c + ph.n_y, c + ph.n_y + ph.n_p))
            self.phase_q_fnc_slices.append(
                slice(c + ph.n_y + ph.n_p, c + ph.n_y + ph.n_p + ph.n_q))
            self.phase_c_slices.append(slice(c, c + ph.n_y + ph.n_p + ph.n_q))
            c += ph.n_y + ph.n_p + ph.n_q
        self.s_var_slice = slice(v, v + self.ir.n_s)
        self.c_endpoint_slice = slice(c, c + self.ir.n_b)

    # -- symbol primitives of the backend (backend.py:81-124, 1343-1384) ----
    # The CUDA backend is sympy-native: the user's symbols ARE its symbols, so the
    # substitution that the CasADi backend performs (sympy -> SX through
    # `user_to_backend_mapping`) reduces to resolving auxiliary data and folded
    # constants -- exactly what the expressions handed to the code generator went through.
    @staticmethod
    def sym(name, rows=1, cols=1):
        import sympy
        if rows == 1 and cols == 1:
            return sympy.Symbol(name)
        return sympy.Matrix(rows, cols, lambda i, j: sympy.Symbol(f"{name}_{i}_{j}"))

    @staticmethod
    def const(val):
        import sympy
        return sympy.Float(val)

    def substitute_pycollo_sym(self, expr, phase=None):
        """Expression over the backend's variables: auxiliary data resolved (with the
        phase's shadowing when ``phase`` -- an index or a phase IR -- is given) and
        constant variables folded (``backend.py:1351-1380``)."""
        import sympy
        if isinstance(expr, (list, tuple)):
            return type(expr)(self.substitute_pycollo_sym(e, phase) for e in expr)
        if isinstance(expr, sympy.MatrixBase):
            return expr.applyfunc(lambda e: self.substitute_pycollo_sym(e, phase))
        if phase is None:
            return self.ir.lower_point(expr, f"'{expr}'")
        ph = self.p[phase] if isinstance(phase, (int, np.integer)) else phase
        return ph.lower(expr, f"'{expr}'")

    @staticmethod
    def expr_as_numeric(expr):
        """``backend.py:1382-1384``: a closed expression as float64."""
        import sympy
        if isinstance(expr, sympy.MatrixBase):
            return np.array(expr.evalf(), dtype=np.float64)
        return np.float64(sympy.sympify(expr).evalf())

    # -- construction steps (optimal_control_problem.py:316-337) ----------
    def create_bounds(self):
        self.bounds = Bounds(self.ir)

    def create_scaling(self):
        self.scaling = Scaling(self)

    def create_quadrature(self):
        self.quadrature = Quadrature(self.ocp.settings.quadrature_method)

    def postprocess_problem_backend(self):
        pass

    def create_initial_mesh(self):
        s = self.ocp.settings
        self.initial_mesh = Mesh(self.quadrature, [ph.mesh for ph in self.ocp.phases],
                                 s.collocation_points_min, s.collocation_points_max)

    def create_guess(self):
        self.initial_guess = Guess(self)

    def create_mesh_iterations(self):
        self.mesh_iterations = []
        self.new_mesh_iteration(self.initial_mesh, self.initial_guess)

    def new_mesh_iteration(self, mesh, guess, batch=1, device=0):
        it = Iteration(self, len(self.mesh_iterations), mesh, guess, batch, device)
        self.mesh_iterations.append(it)
        self.current_iteration = it
        return it

    def iteration_scaling(self, iteration):
        return IterationScaling(iteration)

    # -- NLP function generation (backend.py:1403-1411, 1681-1693) --------
    def generate_nlp_function_callables(self, iteration):
        self.current_iteration = iteration
        iteration.create_engine()

    def create_nlp_solver(self):
        """No IPOPT in this image: the 'solver' is the cyipopt-style callback
        object a host NLP solver consumes (``pycollo/nlp.py:36-76``)."""
        self.nlp_solver = self.nlp_callbacks()
        return self.nlp_solver

    # -- per-callback API (backend.py:1713-1805) ----------------------------
    def _it(self):
        if self.current_iteration is None:
            raise RuntimeError("no mesh iteration has been created")
        return self.current_iteration

    def evaluate_J(self, x):
        return float(self._it().evaluate(_engine.EVAL_F, x)["f"][0])

    def evaluate_g(self, x):
        return self._it().evaluate(_engine.EVAL_GRAD, x)["grad"][0]

    def evaluate_c(self, x):
        return self._it().evaluate(_engine.EVAL_C, x)["c"][0]

    def evaluate_dy(self, x):
        return self._it().evaluate(_engine.EVAL_DY, x)["dy"][0]

    dy_iter_callable = evaluate_dy

    def evaluate_G_nonzeros(self, x):
        return self._it().evaluate(_engine.EVAL_JAC, x)["jac"][0]

    def evaluate_G_structure(self):
        rows, cols = self._it().S.G_structure()
        return rows.copy(), cols.copy()

    def evaluate_G_num_nonzero(self):
        return int(self._it().S.nnz_g)

    def evaluate_G(self, x):
        it = self._it()
        rows, cols = it.S.G_structure()
        return sparse.coo_matrix((self.evaluate_G_nonzeros(x), (rows, cols)),
                                 shape=(it.S.num_c, it.S.num_x))

    def evaluate_H_nonzeros(self, x, obj_factor=1.0, lagrange=None):
        it = self._it()
        if lagrange is None:
            lagrange = np.zeros(it.S.num_c)
        return it.evaluate(_engine.EVAL_HESS, x, lagrange, obj_factor)["hess"][0]

    def _need_second_derivatives(self):
        if int(self.ocp.settings.derivative_level) < 2:
            raise ValueError("derivative_level=1: no Hessian has been generated")

    def evaluate_H_structure(self):
        self._need_second_derivatives()
        rows, cols = self._it().S.H_structure()
        return rows.copy(), cols.copy()

    def evaluate_H_num_nonzero(self):
        self._need_second_derivatives()
        return int(self._it().S.nnz_h)

    def evaluate_H(self, x, obj_factor=1.0, lagrange=None):
        it = self._it()
        rows, cols = it.S.H_structure()
        return sparse.coo_matrix(
            (self.evaluate_H_nonzeros(x, obj_factor, lagrange), (rows, cols)),
            shape=(it.S.num_x, it.S.num_x))

    def nlp_callbacks(self, ordering="cyipopt", x_check="full", register_inputs=True):
        """cyipopt-style callback object (``pycollo/nlp.py:36-76``); without
        ``hessian`` members when ``settings.derivative_level == 1``."""
        from .nlp import NlpCallbacks, NlpCallbacksFirstOrder
        cls = NlpCallbacks if int(self.ocp.settings.derivative_level) >= 2 \
            else NlpCallbacksFirstOrder
        return cls(self._it(), ordering, x_check, register_inputs)

    # -- the callables IterationScaling consumes (scaling.py:361-362, 392-394) ----
    def g_iter_scale_callable(self, x_and_w):
        """``[x_tilde; w_J] -> gradient`` with the objective scale taken from the
        argument instead of the engine's current scaling (``backend.py:1506-1511``): the
        gradient is linear in ``w_J``."""
        xw = np.asarray(x_and_w, dtype=np.float64).ravel()
        it = self._it()
        x, w_arg = xw[:it.S.num_x], xw[it.S.num_x:]
        g = self.evaluate_g(x)
        w_now = float(getattr(getattr(it, "scaling", None), "w", 1.0))
        return g * (float(w_arg[0]) / w_now) if w_arg.size and w_now != 0.0 else g

    def G_iter_scale_callable(self, x_and_W):
        """``[x_tilde; W] -> G`` (scipy COO, CCS entry order) with the OCP-level
        constraint scales taken from the argument (``backend.py:1674-1679``): row i of G
        is linear in its constraint's scale."""
        xw = np.asarray(x_and_W, dtype=np.float64).ravel()
        it = self._it()
        S = it.S
        x, W_arg = xw[:S.num_x], xw[S.num_x:]
        vals = self.evaluate_G_nonzeros(x)
        rows, cols = S.G_structure()
        sc = getattr(it, "scaling", None)
        if W_arg.size and sc is not None:
            ratio = sc._expand_c_to_mesh(W_arg) / sc._expand_c_to_mesh(sc.W_ocp)
            vals = vals * ratio[rows]
        return sparse.coo_matrix((vals, (rows, cols)), shape=(S.num_c, S.num_x))

    def evaluate_dy_on_mesh(self, mesh, x_user, batch=1, device=0):
        """State derivatives ``f(x)`` at every node of ANOTHER mesh of the same problem,
        in the user basis -- the hook that replaces the SX-coupled attributes
        ``mesh_refinement.py:109, 148-151`` reads (``p[i].y_eqn``, ``V/r_sym_val_mapping``)."""
        low = lower_problem(self.ocp, mesh.p)
        eng = _engine.Engine(low.S, low.layouts, low.header, batch=batch, device=device)
        # user basis: V = 1, r = 0, W = 1, w = 1 (mesh_refinement.py:149-150)
        eng.set_scaling(np.ones(low.S.n_var_ocp), np.zeros(low.S.n_var_ocp),
                        np.ones(low.S.n_con_ocp), 1.0)
        dy = eng.eval_host(_engine.EVAL_DY, x_user)["dy"]
        return dy[0] if batch == 1 else dy

    def process_solution(self, iteration, nlp_result):
        """``backend.py:1830-1841``: the processed solution of an NLP result
        (``NlpResult.solution["x"]``, optionally ``["f"]``)."""
        from .solution import Solution
        sol = nlp_result.solution if hasattr(nlp_result, "solution") else nlp_result
        x = np.asarray(sol["x"], dtype=np.float64).ravel()
        J = float(np.asarray(sol["f"]).ravel()[0]) if "f" in sol else None
        return Solution(iteration, x, J)

    def solve_nlp(self, callbacks=None, verbose=False):
        """``backend.py:1807-1827``: solve the NLP of the current mesh iteration from its
        scaled guess inside its scaled bounds; returns ``NlpResult(solution, info,
        solve_time)`` with ``solution`` keyed like ``ca.nlpsol``'s output (``"x"``, ``"f"``,
        ``"g"``, ``"lam_g"``, scaled basis).

        The reference hands the callbacks to IPOPT.  Neither IPOPT nor cyipopt exists in
        this image (attaching cyipopt: INTEGRATION.md), so the host solver is the built-in
        interior-point Newton method (``ipnewton.py``) over the same cyipopt-style callback
        object -- host linear algebra, as IPOPT's is; every value and derivative comes from
        the CUDA callbacks.  ``callbacks``: another object with the same contract (tests)."""
        from . import ipnewton
        it = self._it()
        cb = callbacks if callbacks is not None else self.nlp_callbacks("cyipopt")
        settings = self.ocp.settings
        t0 = timer()
        res = ipnewton.solve(cb, it.guess_x_tilde, it.x_bnd_l, it.x_bnd_u, it.c_bnd_l,
                             it.c_bnd_u, tol=float(settings.nlp_tolerance),
                             max_iter=int(settings.max_nlp_iterations), verbose=verbose)
        solve_time = timer() - t0
        solution = {"x": res.x, "f": res.fun, "g": np.asarray(cb.constraints(res.x)),
                    "lam_g": res.lam}
        info = {"solver": "pycollo_b200.ipnewton", "success": res.success,
                "iterations": res.nit, "kkt_error": res.kkt_error,
                "constraint_violation": res.constr_violation}
        return NlpResult(solution=solution, info=info, solve_time=solve_time)

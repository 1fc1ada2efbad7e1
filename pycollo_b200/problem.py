"""User-facing problem description: the slice of pycollo's API the callback path needs.

This is the host-side *mirror of the reference interface* for the hot path
(task brief, section 2): the same names, argument meaning and error behaviour as
``pycollo.OptimalControlProblem`` / ``Phase`` / ``PhaseBounds`` / ``PhaseGuess`` /
``Settings`` for everything that feeds the NLP callbacks, so that problems are
written exactly as in the reference's ``examples/`` and ``tests/unit/conftest.py``.
It is deliberately thin: validation of exotic inputs, plotting, solution
post-processing and the mesh-refinement heuristics are out of scope
(SURVEY.md §2 rows 13-17 are "OUT OF SCOPE - API layer").

Reference anchors: ``pycollo/optimal_control_problem.py:316-337`` (initialise
order), ``pycollo/phase.py:371-540`` (symbol naming: ``t0_P0``, ``x_P0(t0)``,
``q0_P0``), ``pycollo/bounds.py:640-880`` (bound parsing, constant-variable
detection), ``pycollo/guess.py:124-200`` (guess layout),
``pycollo/settings.py:42-76`` (defaults).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import sympy as sym

from .mesh import PhaseMesh
from .quadrature import (DEFAULT_COLLOCATION_POINTS_MAX,
                         DEFAULT_COLLOCATION_POINTS_MIN, GAUSS, LOBATTO, RADAU)

CUDA = "cuda"
_REFERENCE_BACKENDS = ("casadi", "hsad", "pycollo", "sympy")


class SymbolTuple(tuple):
    """Tuple of symbols/expressions with attribute access by user symbol name
    (the behaviour of the reference's named data containers,
    ``pycollo/utils.py:145-312``)."""

    def __new__(cls, items, names=()):
        self = super().__new__(cls, items)
        self._names = tuple(str(n) for n in names)
        return self

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        try:
            return self[self._names.index(name)]
        except ValueError:
            raise AttributeError(name) from None


def _as_tuple(value):
    if value is None:
        return ()
    if isinstance(value, (sym.Basic, int, float, np.floating, np.integer)):
        return (value,)
    if isinstance(value, dict):
        return tuple(value.values())
    return tuple(value)


class Settings:
    """Validated settings (``pycollo/settings.py``); only keys the path reads."""

    _SCALING = ("bounds", "none", None)

    def __init__(self, ocp=None, **kwargs):
        self.ocp = ocp
        self.backend = CUDA
        self.quadrature_method = LOBATTO
        self.collocation_points_min = DEFAULT_COLLOCATION_POINTS_MIN
        self.collocation_points_max = DEFAULT_COLLOCATION_POINTS_MAX
        self.scaling_method = "bounds"
        self.update_scaling = False
        self.scaling_weight = 0.8
        self.derivative_level = 2
        self.nlp_solver = "ipopt"
        self.linear_solver = "mumps"
        self.nlp_tolerance = 1e-10
        self.max_nlp_iterations = 2000
        self.warm_start = False
        self.mesh_tolerance = 1e-7
        self.max_mesh_iterations = 10
        self.default_number_mesh_sections = 10
        self.default_mesh_section_sizes = None
        self.bound_clash_absolute_tolerance = 1e-6
        self.bound_clash_relative_tolerance = 1e-6
        self.numerical_inf = 10e19       # sic: 1e20, pycollo/bounds.py:35 (its docstring says 1e19)
        self.override_endpoint_bounds = True
        self.remove_constant_variables = True
        self.console_out_progress = False
        self.display_mesh_result_graph = False
        self.display_mesh_result_info = False
        self.display_mesh_refinement_info = False
        self.check_nlp_functions = False
        # engine-specific (new; no reference counterpart)
        self.prune_zero_quadrature_coefficients = True
        self.defer_engine = False        # True: initialise() stops before the device
        for key, value in kwargs.items():
            setattr(self, key, value)

    def __setattr__(self, key, value):
        if key == "backend":
            value = str(value).lower()
            if value in _REFERENCE_BACKENDS:
                raise ValueError(
                    f"'{value}' is not available in pycollo_b200: this package "
                    f"provides only the '{CUDA}' backend (the reference's "
                    f"CasADi path is the thing it replaces).")
            if value != CUDA:
                raise ValueError(f"'{value}' is not a valid option of Pycollo "
                                 f"backend. Choose one of: '{CUDA}'.")
        elif key == "quadrature_method":
            value = str(value).lower()
            if value == GAUSS:
                raise ValueError(f"'{GAUSS}' is not currently supported as a "
                                 f"quadrature method.")
            if value not in (LOBATTO, RADAU):
                raise ValueError(f"'{value}' is not a valid quadrature method.")
        elif key == "scaling_method":
            if value not in self._SCALING:
                if value in ("guess", "user"):
                    raise ValueError(f"'{value}' is not currently supported as "
                                     f"a scaling method.")
                raise ValueError(f"'{value}' is not a valid scaling method.")
        elif key == "derivative_level":
            value = int(value)
            if value not in (1, 2):
                raise ValueError("derivative_level must be 1 or 2.")
        elif key in ("collocation_points_min", "collocation_points_max"):
            value = int(value)
            if not 2 <= value <= 20:
                raise ValueError(f"{key} must be between 2 and 20.")
        object.__setattr__(self, key, value)


class PhaseBounds:
    """Bounds of one phase (``pycollo/bounds.py:118-330``)."""

    def __init__(self, phase):
        self.phase = phase
        self.initial_time = None
        self.final_time = None
        self.state_variables = None
        self.control_variables = None
        self.integral_variables = None
        self.path_constraints = None
        self.initial_state_constraints = None
        self.final_state_constraints = None


class EndpointBounds:
    """Problem-level bounds (``pycollo/bounds.py:36-115``)."""

    def __init__(self, ocp):
        self.ocp = ocp
        self.parameter_variables = None
        self.endpoint_constraints = None


class PhaseGuess:
    """Initial guess of one phase (``pycollo/guess.py:12-70``)."""

    def __init__(self, phase):
        self.phase = phase
        self.time = None
        self.state_variables = None
        self.control_variables = None
        self.integral_variables = None


class EndpointGuess:
    def __init__(self, ocp):
        self.ocp = ocp
        self.parameter_variables = None


class Phase:
    """One phase of the OCP (``pycollo/phase.py``)."""

    def __init__(self, name, optimal_control_problem, *, state_variables=None,
                 control_variables=None, state_equations=None,
                 integrand_functions=None, path_constraints=None,
                 auxiliary_data=None):
        self.name = name
        self.optimal_control_problem = optimal_control_problem
        self.phase_number = len(optimal_control_problem._phases)
        self._suffix = str(self.phase_number)
        optimal_control_problem._phases.append(self)
        self._t0_USER = sym.Symbol(f"t0_P{self._suffix}")
        self._tF_USER = sym.Symbol(f"tF_P{self._suffix}")
        self.state_variables = state_variables
        self.control_variables = control_variables
        self.state_equations = state_equations
        self.integrand_functions = integrand_functions
        self.path_constraints = path_constraints
        self.auxiliary_data = dict(auxiliary_data) if auxiliary_data else {}
        self.bounds = PhaseBounds(self)
        self.guess = PhaseGuess(self)
        s = optimal_control_problem.settings
        self.mesh = PhaseMesh(s.default_number_mesh_sections,
                              s.default_mesh_section_sizes,
                              s.collocation_points_min)

    # -- variables ------------------------------------------------------
    @property
    def state_variables(self):
        return self._y_var_user

    @state_variables.setter
    def state_variables(self, y_vars):
        y = _as_tuple(y_vars)
        names = [str(v) for v in y]
        if len(set(names)) != len(names):
            raise ValueError("State variable names must be unique.")
        self._y_var_user = SymbolTuple(y, names)
        sfx = self._suffix
        self._y_t0_user = SymbolTuple(
            [sym.Symbol(f"{n}_P{sfx}(t0)") for n in names], names)
        self._y_tF_user = SymbolTuple(
            [sym.Symbol(f"{n}_P{sfx}(tF)") for n in names], names)

    @property
    def control_variables(self):
        return self._u_var_user

    @control_variables.setter
    def control_variables(self, u_vars):
        u = _as_tuple(u_vars)
        self._u_var_user = SymbolTuple(u, [str(v) for v in u])

    @property
    def initial_state_variables(self):
        return self._y_t0_user

    @property
    def final_state_variables(self):
        return self._y_tF_user

    @property
    def initial_time_variable(self):
        return self._t0_USER

    @property
    def final_time_variable(self):
        return self._tF_USER

    @property
    def time_variables(self):
        return (self._t0_USER, self._tF_USER)

    @property
    def integral_variables(self):
        return self._q_var_user

    # -- equations ------------------------------------------------------
    @property
    def state_equations(self):
        return self._y_eqn_user

    @state_equations.setter
    def state_equations(self, y_eqns):
        if isinstance(y_eqns, dict):
            order = {str(v): i for i, v in enumerate(self._y_var_user)}
            items = sorted(y_eqns.items(), key=lambda kv: order[str(kv[0])])
            y_eqns = [v for _, v in items]
        self._y_eqn_user = tuple(sym.sympify(e) for e in _as_tuple(y_eqns))

    @property
    def integrand_functions(self):
        return self._q_fnc_user

    @integrand_functions.setter
    def integrand_functions(self, integrands):
        self._q_fnc_user = tuple(sym.sympify(e) for e in _as_tuple(integrands))
        self._q_var_user = SymbolTuple(
            [sym.Symbol(f"q{i}_P{self._suffix}")
             for i in range(len(self._q_fnc_user))],
            [f"q{i}" for i in range(len(self._q_fnc_user))])

    @property
    def path_constraints(self):
        return self._p_con_user

    @path_constraints.setter
    def path_constraints(self, p_cons):
        self._p_con_user = tuple(sym.sympify(e) for e in _as_tuple(p_cons))

    number_state_variables = property(lambda self: len(self._y_var_user))
    number_control_variables = property(lambda self: len(self._u_var_user))
    number_integral_variables = property(lambda self: len(self._q_var_user))
    number_state_equations = property(lambda self: len(self._y_eqn_user))
    number_path_constraints = property(lambda self: len(self._p_con_user))
    number_integrand_functions = property(lambda self: len(self._q_fnc_user))

    def _check_variables_and_equations(self):
        """``pycollo/phase.py:571-661``: one state equation per state."""
        if len(self._y_eqn_user) != len(self._y_var_user):
            raise ValueError(
                f"A state equation must be supplied for each state variable in "
                f"phase {self.name}: {len(self._y_var_user)} state variables "
                f"but {len(self._y_eqn_user)} state equations.")


class OptimalControlProblem:
    """The OCP container (``pycollo/optimal_control_problem.py``)."""

    def __init__(self, name="Untitled", parameter_variables=None, *,
                 objective_function=None, endpoint_constraints=None,
                 auxiliary_data=None, settings=None):
        self.name = name
        self.settings = settings if settings is not None else Settings(self)
        self.settings.ocp = self
        self._phases = []
        self.parameter_variables = parameter_variables
        self.objective_function = objective_function
        self.endpoint_constraints = endpoint_constraints
        self.auxiliary_data = dict(auxiliary_data) if auxiliary_data else {}
        self.bounds = EndpointBounds(self)
        self.guess = EndpointGuess(self)
        self._backend = None
        self._is_initialised = False

    @property
    def phases(self):
        return SymbolTuple(self._phases, [p.name for p in self._phases])

    @property
    def number_phases(self):
        return len(self._phases)

    def new_phase(self, name, **kwargs):
        return Phase(name, self, **kwargs)

    def new_phase_like(self, phase_for_copying, name, **kwargs):
        return self.new_phases_like(phase_for_copying, 1, [name], **kwargs)[0]

    def new_phases_like(self, phase_for_copying, number, names, *,
                        copy_state_variables=True, copy_control_variables=True,
                        copy_state_equations=True, copy_path_constraints=True,
                        copy_integrand_functions=True,
                        copy_state_endpoint_constraints=False,
                        copy_bounds=True, copy_mesh=True, copy_scaling=True,
                        copy_guess=True):
        """``pycollo/phase.py:156-218`` semantics for the fields modelled here."""
        import copy
        src = phase_for_copying
        out = []
        for name in list(names)[:number]:
            ph = Phase(name, self)
            if copy_state_variables:
                ph.state_variables = src.state_variables
                if copy_bounds:
                    ph.bounds.state_variables = copy.deepcopy(
                        src.bounds.state_variables)
                if copy_guess:
                    ph.guess.state_variables = copy.deepcopy(
                        src.guess.state_variables)
            if copy_control_variables:
                ph.control_variables = src.control_variables
                if copy_bounds:
                    ph.bounds.control_variables = copy.deepcopy(
                        src.bounds.control_variables)
                if copy_guess:
                    ph.guess.control_variables = copy.deepcopy(
                        src.guess.control_variables)
            if copy_state_equations:
                ph.state_equations = src.state_equations
            if copy_path_constraints:
                ph.path_constraints = src.path_constraints
                if copy_bounds:
                    ph.bounds.path_constraints = copy.deepcopy(
                        src.bounds.path_constraints)
            if copy_integrand_functions:
                ph.integrand_functions = src.integrand_functions
                if copy_bounds:
                    ph.bounds.integral_variables = copy.deepcopy(
                        src.bounds.integral_variables)
                if copy_guess:
                    ph.guess.integral_variables = copy.deepcopy(
                        src.guess.integral_variables)
            if copy_state_endpoint_constraints and copy_bounds:
                ph.bounds.initial_state_constraints = copy.deepcopy(
                    src.bounds.initial_state_constraints)
                ph.bounds.final_state_constraints = copy.deepcopy(
                    src.bounds.final_state_constraints)
            if copy_mesh:
                ph.mesh = copy.deepcopy(src.mesh)
            out.append(ph)
        return out

    @property
    def parameter_variables(self):
        return self._s_var_user

    @parameter_variables.setter
    def parameter_variables(self, s_vars):
        s = _as_tuple(s_vars)
        self._s_var_user = SymbolTuple(s, [str(v) for v in s])

    @property
    def endpoint_constraints(self):
        return self._b_con_user

    @endpoint_constraints.setter
    def endpoint_constraints(self, b_cons):
        self._b_con_user = tuple(sym.sympify(e) for e in _as_tuple(b_cons))

    @property
    def objective_function(self):
        return self._J_user

    @objective_function.setter
    def objective_function(self, J):
        self._J_user = None if J is None else sym.sympify(J)

    number_parameter_variables = property(lambda self: len(self._s_var_user))
    number_endpoint_constraints = property(lambda self: len(self._b_con_user))

    # -- lifecycle (``optimal_control_problem.py:316-337``) ---------------
    def initialise(self):
        from .backend import Cuda
        for p in self._phases:
            p._check_variables_and_equations()
        if self._J_user is None:
            raise ValueError("An objective function must be supplied.")
        self._backend = Cuda(self)
        self._backend.create_bounds()
        self._backend.create_scaling()
        self._backend.create_quadrature()
        self._backend.postprocess_problem_backend()
        self._backend.create_initial_mesh()
        self._backend.create_guess()
        self._backend.create_mesh_iterations()
        self._is_initialised = True
        return self

    def solve(self, display_progress=False):
        """ONE mesh iteration of the reference's solve (``optimal_control_problem.py:
        339-...``, ``iteration.py:474-526``): NLP solve from the guess -> processed
        solution (``casadi_solution.py``) -> mesh-refinement error of the solved mesh
        (``mesh_refinement.py:60-73``); sets ``nlp_result``, ``solution``,
        ``mesh_refinement`` and ``mesh_tolerance_met`` and returns the solution.

        The host NLP solver is the backend's (``Cuda.solve_nlp``: the built-in
        interior-point method; IPOPT / cyipopt are absent from this image).  Building the
        NEXT mesh from the errors and looping (Patterson-Rao re-meshing) is the
        reference's control plane, outside this package's scope:
        a caller that has it continues with ``_backend.new_mesh_iteration(mesh, guess)``.
        """
        if not getattr(self, "_is_initialised", False):
            self.initialise()
        backend = self._backend
        iteration = backend.current_iteration
        self.nlp_result = backend.solve_nlp(verbose=display_progress)
        self.solution = backend.process_solution(iteration, self.nlp_result)
        self.mesh_refinement = self.solution.refine_mesh()
        worst = max([float(np.max(m)) for m in
                     self.mesh_refinement.maximum_relative_mesh_errors if np.size(m)] + [0.0])
        self.mesh_tolerance_met = bool(worst <= float(self.settings.mesh_tolerance))
        if display_progress:
            info = self.nlp_result.info
            print(f"mesh iteration 0: J = {self.solution.objective:.12g} in "
                  f"{info['iterations']} NLP iterations ({self.nlp_result.solve_time:.2f} s), "
                  f"largest relative mesh error {worst:.3e} "
                  f"(tolerance {float(self.settings.mesh_tolerance):.1e})")
        return self.solution

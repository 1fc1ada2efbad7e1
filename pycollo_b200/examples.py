"""Problem definitions used by the tests, ``bench.py`` and ``smoke()``.

Each builder returns a fully defined, un-initialised ``OptimalControlProblem``
written exactly as a pycollo user would write it.  The problem data
(equations, bounds, guesses) are those of the reference's example scripts and
unit-test fixtures, cited per function; they are *inputs* to the engine.

``set_mesh`` scales a problem to BASELINE.json's mesh sizes.
"""
from __future__ import annotations

import numpy as np
import sympy as sym

from .mesh import PhaseMesh
from .problem import OptimalControlProblem


def set_mesh(problem, number_mesh_sections, number_mesh_section_nodes=4,
             mesh_section_sizes=None, phases=None):
    """Give (selected) phases a mesh of K sections x N_k nodes."""
    for ph in (problem.phases if phases is None else phases):
        ph.mesh = PhaseMesh(number_mesh_sections, mesh_section_sizes,
                            number_mesh_section_nodes)
    return problem


def brachistochrone(quadrature_method="lobatto", scaling_method="bounds"):
    """``examples/brachistochrone/brachistochrone.py``,
    ``tests/unit/conftest.py:14-76`` (n_y=3, n_u=1, n_t=1)."""
    x, y, v, u = sym.symbols("x y v u")
    g = 9.81
    problem = OptimalControlProblem(name="Brachistochrone")
    phase = problem.new_phase(name="A")
    phase.state_variables = [x, y, v]
    phase.control_variables = u
    phase.state_equations = [v * sym.sin(u), v * sym.cos(u), g * sym.cos(u)]
    phase.auxiliary_data = {}
    problem.objective_function = phase.final_time_variable
    phase.bounds.initial_time = 0.0
    phase.bounds.final_time = [0, 10]
    phase.bounds.state_variables = [[0, 10], [0, 10], [-50, 50]]
    phase.bounds.control_variables = [[-np.pi / 2, np.pi / 2]]
    phase.bounds.initial_state_constraints = {x: 0, y: 0, v: 0}
    phase.bounds.final_state_constraints = {x: 2, y: 2}
    phase.guess.time = np.array([0, 10])
    phase.guess.state_variables = np.array([[0, 2], [0, 2], [0, 0]])
    phase.guess.control_variables = np.array([[0, np.pi / 2]])
    problem.settings.derivative_level = 2
    problem.settings.scaling_method = scaling_method
    problem.settings.quadrature_method = quadrature_method
    return problem


def double_pendulum():
    """``tests/unit/conftest.py:79-190`` (aux-data heavy; n_y=4, n_u=2, n_q=1,
    n_t=1, n_s=2; phase-level aux data shadow problem-level ones)."""
    a0, a1, v0, v1, T0, T1 = sym.symbols("a0 a1 v0 v1 T0 T1")
    g = sym.symbols("g")
    m0, p0, d0, l0, k0, I0 = sym.symbols("m0 p0 d0 l0 k0 I0")
    m1, p1, d1, l1, k1, I1 = sym.symbols("m1 p1 d1 l1 k1 I1")
    c0, s0, c1, s1 = sym.symbols("c0 s0 c1 s1")
    M00, M01, M10, M11, K0, K1 = sym.symbols("M00 M01 M10 M11 K0 K1")
    detM = sym.symbols("detM")
    K0_eqn = (T0 + g * (m0 * p0 + m1 * l0) * c0
              + m1 * p1 * l0 * (s1 * c0 - s0 * c1) * v1 ** 2)
    K1_eqn = (T1 + g * m1 * p1 * c1
              + m1 * p1 * l0 * (s0 * c1 - s1 * c0) * v0 ** 2)
    problem = OptimalControlProblem(name="Double Pendulum Swing-Up")
    phase = problem.new_phase(name="A")
    phase.state_variables = [a0, a1, v0, v1]
    phase.control_variables = [T0, T1]
    phase.state_equations = [v0, v1, (M11 * K0 - M01 * K1) / detM,
                             (M00 * K1 - M10 * K0) / detM]
    phase.integrand_functions = [(T0 ** 2 + T1 ** 2)]
    phase.auxiliary_data = {g: -9.81, k1: 1 / 12,
                            I0: m0 * (k0 ** 2 + p0 ** 2),
                            I1: m1 * (k1 ** 2 + p1 ** 2),
                            s0: sym.sin(a0), c1: sym.cos(a1)}
    problem.parameter_variables = [m0, p0]
    problem.objective_function = phase.integral_variables[0]
    problem.auxiliary_data = {
        g: 0, d0: 0.5, k0: 1 / 12, m1: 1.0, p1: 0.5, d1: 0.5,
        l0: p0 + d0, l1: p1 + d1,
        I0: m0 * (k0 ** 2 + p0 ** 2), I1: m1 * (k1 ** 2 + p1 ** 2),
        c0: sym.cos(a0), s0: sym.sin(a0), c1: sym.cos(a1), s1: sym.sin(a1),
        M00: I0 + m1 * l0 ** 2, M01: m1 * p1 * l0 * (s0 * s1 + c0 * c1),
        M10: M01, M11: I1, K0: K0_eqn, K1: K1_eqn,
        detM: M00 * M11 - M01 * M10}
    phase.bounds.initial_time = 0
    phase.bounds.final_time = [1, 3]
    phase.bounds.state_variables = [[-np.pi, np.pi], [-np.pi, np.pi],
                                    [-10, 10], [-10, 10]]
    phase.bounds.control_variables = [[-15, 15], [-15, 15]]
    phase.bounds.integral_variables = [0, 1000]
    phase.bounds.initial_state_constraints = [
        [-0.5 * np.pi, -0.5 * np.pi], [-0.5 * np.pi, -0.5 * np.pi],
        [0, 0], [0, 0]]
    phase.bounds.final_state_constraints = [
        [0.5 * np.pi, 0.5 * np.pi], [0.5 * np.pi, 0.5 * np.pi], [0, 0], [0, 0]]
    problem.bounds.parameter_variables = [[0.5, 1.5], [0.5, 1.5]]
    phase.guess.time = [0, 2]
    phase.guess.state_variables = [[-0.5 * np.pi, 0.5 * np.pi],
                                   [-0.5 * np.pi, 0.5 * np.pi], [0, 0], [0, 0]]
    phase.guess.control_variables = [[0, 0], [0, 0]]
    phase.guess.integral_variables = [[100]]
    problem.guess.parameter_variables = [1.0, 1.0]
    return problem


def hypersensitive(quadrature_method="lobatto"):
    """``examples/hypersensitive_problem/hypersensitive_problem.py``
    (n_y = n_u = n_q = 1, fixed times)."""
    y, u = sym.symbols("y u")
    problem = OptimalControlProblem(name="Hypersensitive problem")
    phase = problem.new_phase(name="A")
    phase.state_variables = y
    phase.control_variables = u
    phase.state_equations = [-y ** 3 + u]
    phase.integrand_functions = [0.5 * (y ** 2 + u ** 2)]
    phase.auxiliary_data = {}
    phase.bounds.initial_time = 0.0
    phase.bounds.final_time = 10000.0
    phase.bounds.state_variables = [[-50, 50]]
    phase.bounds.control_variables = [[-50, 50]]
    phase.bounds.integral_variables = [[0, 100000]]
    phase.bounds.initial_state_constraints = [[1.0, 1.0]]
    phase.bounds.final_state_constraints = [[1.5, 1.5]]
    phase.guess.time = np.array([0.0, 10000.0])
    phase.guess.state_variables = np.array([[1.0, 1.5]])
    phase.guess.control_variables = np.array([[0.0, 0.0]])
    phase.guess.integral_variables = np.array([4])
    problem.objective_function = phase.integral_variables[0]
    problem.settings.quadrature_method = quadrature_method
    return problem


def cart_pole_swing_up(quadrature_method="lobatto", scaling_method="bounds"):
    """``examples/cart_pole_swing_up/cart_pole_swing_up_explicit.py``
    (n_y=4, n_u=1, n_q=1, n_t=0) -- the BASELINE.json config-2 problem."""
    q1, q2, q1d, q2d, q1dd, q2dd, F = sym.symbols("q1 q2 q1d q2d q1dd q2dd F")
    m1, m2, l, g = sym.symbols("m1 m2 l g")
    F_max, d_max, d, T = 20.0, 2.0, 1.0, 2.0
    problem = OptimalControlProblem(name="Cart-Pole Swing-Up")
    phase = problem.new_phase(name="A")
    phase.state_variables = [q1, q2, q1d, q2d]
    phase.control_variables = F
    phase.state_equations = [q1d, q2d, q1dd, q2dd]
    phase.integrand_functions = [F ** 2]
    phase.bounds.initial_time = 0
    phase.bounds.final_time = T
    phase.bounds.state_variables = {q1: [-d_max, d_max], q2: [-10, 10],
                                    q1d: [-10, 10], q2d: [-10, 10]}
    phase.bounds.control_variables = {F: [-F_max, F_max]}
    phase.bounds.integral_variables = [[0, 100]]
    phase.bounds.initial_state_constraints = {q1: 0, q2: 0, q1d: 0, q2d: 0}
    phase.bounds.final_state_constraints = {q1: d, q2: np.pi, q1d: 0, q2d: 0}
    phase.guess.time = [0, T]
    phase.guess.state_variables = [[0, d], [0, np.pi], [0, 0], [0, 0]]
    phase.guess.control_variables = [[0, 0]]
    phase.guess.integral_variables = [0]
    q1dd_eqn = (l * m2 * sym.sin(q2) * q2d ** 2 + F
                + m2 * g * sym.cos(q2) * sym.sin(q2)) \
        / (m1 + m2 * (1 - sym.cos(q2) ** 2))
    q2dd_eqn = -(l * m2 * sym.cos(q2) * sym.sin(q2) * q2d ** 2
                 + F * sym.cos(q2) + (m1 + m2) * g * sym.sin(q2)) \
        / (l * m1 + l * m2 * (1 - sym.cos(q2) ** 2))
    problem.objective_function = phase.integral_variables[0]
    problem.auxiliary_data = {g: 9.81, l: 0.5, m1: 1.0, m2: 0.3,
                              q1dd: q1dd_eqn, q2dd: q2dd_eqn}
    problem.settings.quadrature_method = quadrature_method
    problem.settings.scaling_method = scaling_method
    return problem

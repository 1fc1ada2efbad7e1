"""User scripts: the optimal control problems the tests, ``bench.py``, ``smoke()``
and the golden-vector generator evaluate.

Each builder returns a fully defined, un-initialised ``OptimalControlProblem``
written exactly as a pycollo user writes it -- so exactly that the same function
builds the problem on EITHER package: ``api=None`` (default) uses
``pycollo_b200``; ``oracle/make_golden_nlp.py`` passes the reference's own
``pycollo`` module (imported from ``/root/reference`` in the build container) to
produce the golden vectors the CUDA path is checked against.  The problem data
(equations, bounds, guesses) are those of the reference's example scripts and
unit-test fixtures, cited per function; they are *inputs* to the engine.

``set_mesh`` scales a problem to BASELINE.json's mesh sizes.
"""
from __future__ import annotations

import numpy as np
import sympy as sym


def _ocp_class(api):
    if api is None:
        import pycollo_b200 as api
    return api.OptimalControlProblem


def set_mesh(problem, number_mesh_sections, number_mesh_section_nodes=4,
             mesh_section_sizes=None, phases=None):
    """Give (selected) phases a mesh of K sections x N_k nodes (either package:
    the three user-facing attributes of ``pycollo/mesh.py:10-121``)."""
    for ph in (problem.phases if phases is None else phases):
        m = ph.mesh
        if hasattr(type(m), "number_mesh_sections"):      # reference: validated setters
            m._mesh_sec_sizes = None
            m.number_mesh_sections = number_mesh_sections
            m.mesh_section_sizes = mesh_section_sizes
            m.number_mesh_section_nodes = number_mesh_section_nodes
        else:
            ph.mesh = type(m)(number_mesh_sections, mesh_section_sizes,
                              number_mesh_section_nodes)
    return problem


def brachistochrone(quadrature_method="lobatto", scaling_method="bounds", api=None):
    """``examples/brachistochrone/brachistochrone.py``,
    ``tests/unit/conftest.py:14-76`` (n_y=3, n_u=1, n_t=1)."""
    x, y, v, u = sym.symbols("x y v u")
    g = 9.81
    problem = _ocp_class(api)(name="Brachistochrone")
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


def double_pendulum(api=None):
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
    problem = _ocp_class(api)(name="Double Pendulum Swing-Up")
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


def hypersensitive(quadrature_method="lobatto", api=None):
    """``examples/hypersensitive_problem/hypersensitive_problem.py``
    (n_y = n_u = n_q = 1, fixed times)."""
    y, u = sym.symbols("y u")
    problem = _ocp_class(api)(name="Hypersensitive problem")
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


def cart_pole_swing_up(quadrature_method="lobatto", scaling_method="bounds", api=None):
    """``examples/cart_pole_swing_up/cart_pole_swing_up_explicit.py``
    (n_y=4, n_u=1, n_q=1, n_t=0) -- the BASELINE.json config-2 problem."""
    q1, q2, q1d, q2d, q1dd, q2dd, F = sym.symbols("q1 q2 q1d q2d q1dd q2dd F")
    m1, m2, l, g = sym.symbols("m1 m2 l g")
    F_max, d_max, d, T = 20.0, 2.0, 1.0, 2.0
    problem = _ocp_class(api)(name="Cart-Pole Swing-Up")
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


def free_flying_robot(quadrature_method="lobatto", api=None):
    """``examples/free_flying_robot/free_flying_robot.py`` (n_y=6, n_u=4, n_p=2,
    n_q=1, fixed times) -- BASELINE config 3."""
    r_x, r_y, theta, v_x, v_y, omega = sym.symbols("r_x r_y theta v_x v_y omega")
    u_x_pos, u_x_neg, u_y_pos, u_y_neg = sym.symbols("u_x_pos u_x_neg u_y_pos u_y_neg")
    T_x, T_y, I_xx, I_yy = sym.symbols("T_x T_y I_xx I_yy")
    problem = _ocp_class(api)(name="Free-Flying Robot")
    phase = problem.new_phase(name="A",
                              state_variables=[r_x, r_y, theta, v_x, v_y, omega],
                              control_variables=[u_x_pos, u_x_neg, u_y_pos, u_y_neg])
    phase.state_equations = {r_x: v_x, r_y: v_y, theta: omega,
                             v_x: (T_x + T_y) * sym.cos(theta),
                             v_y: (T_x + T_y) * sym.sin(theta),
                             omega: (I_xx * T_x) - (I_yy * T_y)}
    phase.integrand_functions = [u_x_pos + u_x_neg + u_y_pos + u_y_neg]
    phase.path_constraints = [(u_x_pos + u_x_neg), (u_y_pos + u_y_neg)]
    problem.objective_function = phase.integral_variables[0]
    problem.auxiliary_data = {I_xx: 0.2, I_yy: 0.2, T_x: u_x_pos - u_x_neg,
                              T_y: u_y_pos - u_y_neg}
    phase.bounds.initial_time = 0.0
    phase.bounds.final_time = 12.0
    phase.bounds.state_variables = {r_x: [-10, 10], r_y: [-10, 10],
                                    theta: [-np.pi, np.pi], v_x: [-2, 2],
                                    v_y: [-2, 2], omega: [-1, 1]}
    phase.bounds.initial_state_constraints = {
        r_x: [-10, -10], r_y: [-10, -10], theta: [np.pi / 2, np.pi / 2],
        v_x: [0, 0], v_y: [0, 0], omega: [0, 0]}
    phase.bounds.final_state_constraints = {
        r_x: [0, 0], r_y: [0, 0], theta: [0, 0], v_x: [0, 0], v_y: [0, 0],
        omega: [0, 0]}
    phase.bounds.control_variables = {u_x_pos: [0, 1000], u_x_neg: [0, 1000],
                                      u_y_pos: [0, 1000], u_y_neg: [0, 1000]}
    phase.bounds.integral_variables = [[0, 100]]
    phase.bounds.path_constraints = [[-1000, 1], [-1000, 1]]
    phase.guess.time = [0.0, 12.0]
    phase.guess.state_variables = [[-10, 0], [-10, 0], [np.pi / 2, 0], [0, 0],
                                   [0, 0], [0, 0]]
    phase.guess.control_variables = [[0, 0], [0, 0], [0, 0], [0, 0]]
    phase.guess.integral_variables = [0]
    problem.settings.mesh_tolerance = 1e-7
    problem.settings.max_mesh_iterations = 25
    problem.settings.quadrature_method = quadrature_method
    return problem


def space_shuttle_reentry(quadrature_method="lobatto", api=None):
    """``examples/space_shuttle_reentry_trajectory/...maximum_crossrange.py``
    (n_y=6, n_u=2, free final time) -- BASELINE config 3."""
    h, phi, theta, nu, gamma, psi, alpha, beta = sym.symbols(
        "h phi theta nu gamma psi alpha beta")
    D, L, g, r, rho, rho_0, h_r, c_L, c_D, Re, S = sym.symbols(
        "D L g r rho rho_0 h_r c_L c_D Re S")
    c_lift_0, c_lift_1, mu, c_drag_0, c_drag_1, c_drag_2, m = sym.symbols(
        "c_lift_0 c_lift_1 mu c_drag_0 c_drag_1 c_drag_2 m")
    problem = _ocp_class(api)(
        name="Space shuttle reentry trajectory maximum crossrange")
    phase = problem.new_phase(name="A")
    phase.state_variables = [h, phi, theta, nu, gamma, psi]
    phase.control_variables = [alpha, beta]
    dgamma_1 = L * sym.cos(beta) / (m * nu)
    dgamma_2 = sym.cos(gamma) * ((nu / r) - (g / nu))
    dpsi_1 = L * sym.sin(beta) / (m * nu * sym.cos(gamma))
    dpsi_2 = nu * sym.cos(gamma) * sym.sin(psi) * sym.sin(theta)
    dpsi_3 = r * sym.cos(theta)
    phase.state_equations = {
        h: nu * sym.sin(gamma),
        phi: nu * sym.cos(gamma) * sym.sin(psi) / (r * sym.cos(theta)),
        theta: nu * sym.cos(gamma) * sym.cos(psi) / r,
        nu: -(D / m) - g * sym.sin(gamma),
        gamma: dgamma_1 + dgamma_2,
        psi: dpsi_1 + dpsi_2 / dpsi_3}
    problem.objective_function = -phase.final_state_variables[2]
    problem.auxiliary_data = {
        rho_0: 1.225570827014494, h_r: 7254.24, Re: 6371203.92, S: 249.9091776,
        c_lift_0: -0.2070, c_lift_1: 1.6756, mu: 3.986031954093051e14,
        c_drag_0: 0.07854, c_drag_1: -0.3529, c_drag_2: 2.0400,
        D: 0.5 * c_D * S * rho * nu ** 2, L: 0.5 * c_L * S * rho * nu ** 2,
        g: mu / (r ** 2), r: Re + h, rho: rho_0 * sym.exp(-h / h_r),
        c_L: c_lift_0 + (c_lift_1 * alpha),
        c_D: c_drag_0 + (c_drag_1 * alpha) + (c_drag_2 * alpha ** 2),
        m: 92079.2525560557}
    d2r = np.pi / 180
    phase.bounds.initial_time = [0.0, 0.0]
    phase.bounds.final_time = [0.0, 3000.0]
    phase.bounds.state_variables = {
        h: [0, 300000], phi: [-np.pi, np.pi], theta: [-70 * d2r, 70 * d2r],
        nu: [10, 45000], gamma: [-80 * d2r, 80 * d2r], psi: [-np.pi, np.pi]}
    phase.bounds.control_variables = {alpha: [-np.pi / 2, np.pi / 2],
                                      beta: [-np.pi / 2, np.pi / 180]}
    phase.bounds.initial_state_constraints = {
        h: 79248, phi: 0, theta: 0, nu: 7802.88, gamma: -1 * d2r, psi: 90 * d2r}
    phase.bounds.final_state_constraints = {
        h: [24384, 24384], nu: [762, 762], gamma: [-5 * d2r, -5 * d2r]}
    phase.guess.time = np.array([0.0, 1000])
    phase.guess.state_variables = np.array([
        [79248, 24384], [0, 10 * d2r], [0, 10 * d2r], [7802.88, 762],
        [-1 * d2r, -5 * d2r], [90 * d2r, -90 * d2r]])
    phase.guess.control_variables = np.array([[0, 0], [0, 0]])
    problem.settings.quadrature_method = quadrature_method
    return problem


def multiphase_sliding_mass(num_phases=3, quadrature_method="lobatto", api=None):
    """``tests/integration/test_multiphase.py:25-75``: unit mass slid from 0 to 1,
    split in phases linked by endpoint constraints on velocity and time."""
    x, v, f = sym.symbols("x v f")
    MAX_T, MAX_V, MAX_F = 1.0, 10.0, 20.0
    names = {0: "A", 1: "B", 2: "C", 3: "D"}
    problem = _ocp_class(api)(f"{num_phases}-phase Sliding Mass")
    for i in range(num_phases):
        start_x, end_x = i / num_phases, (i + 1) / num_phases
        phase = problem.new_phase(names[i], state_variables=[x, v],
                                  control_variables=[f])
        phase.state_equations = {x: v, v: f}
        phase.bounds.initial_time = [0, MAX_T] if i else 0
        phase.bounds.final_time = [0, MAX_T]
        phase.bounds.initial_state_constraints = {x: start_x,
                                                  v: [0, MAX_V] if i else 0}
        phase.bounds.state_variables = {x: [start_x, end_x], v: [0, MAX_V]}
        phase.bounds.final_state_constraints = {
            x: end_x, v: [0, MAX_V] if ((i + 1) != num_phases) else 0}
        phase.bounds.control_variables = {f: [-MAX_F, MAX_F]}
        phase.guess.time = [start_x * MAX_T, end_x * MAX_T]
        phase.guess.state_variables = [[start_x, end_x], [0, 0]]
        phase.guess.control_variables = [[0, 0]]
    if num_phases >= 2:
        cons = []
        for p1, p2 in zip(problem.phases[:-1], problem.phases[1:]):
            cons.append(p1.final_state_variables.v - p2.initial_state_variables.v)
            cons.append(p1.final_time_variable - p2.initial_time_variable)
        problem.endpoint_constraints = cons
        problem.bounds.endpoint_constraints = [[0, 0]] * len(cons)
    problem.objective_function = problem.phases[-1].final_time_variable
    problem.settings.quadrature_method = quadrature_method
    return problem


def delta_iii_launch_vehicle(api=None):
    """``examples/delta_iii_launch_vehicle/delta_iii_launch_vehicle.py``: 4 phases,
    n_y=7, n_u=3, n_p=2 each, 18 linkage constraints, phase-dependent auxiliary
    data (T, xi) -- BASELINE config 4.  The reference marks this example as not
    solvable with pycollo (``:3``); it is used for callback parity/throughput."""
    r_x, r_y, r_z, v_x, v_y, v_z, m = sym.symbols("r_x r_y r_z v_x v_y v_z m")
    u_x, u_y, u_z = sym.symbols("u_x u_y u_z")
    D_x, D_y, D_z, T, xi, C_D, S, omega_E = sym.symbols("D_x D_y D_z T xi C_D S omega_E")
    v_r_x, v_r_y, v_r_z = sym.symbols("v_r_x v_r_y v_r_z")
    oxr_x, oxr_y, oxr_z = sym.symbols("omega_x_r_x omega_x_r_y omega_x_r_z")
    mu, R_E, psi_L, g_0, h_0, h, rho, rho_0 = sym.symbols("mu R_E psi_L g_0 h_0 h rho rho_0")
    r_vec_norm, u_vec_norm, v_vec_norm, v_r_vec_norm, T_over_m = sym.symbols(
        "r_vec_norm u_vec_norm v_vec_norm v_r_vec_norm T_over_m")
    t_launch, t_sep_S, t_sep_1, t_sep_2, t_orbit = 0.0, 75.2, 150.4, 261, 961
    m_tot_S, m_tot_1, m_tot_2 = 19290, 104380, 19300
    m_prop_S, m_prop_1 = 17010, 95550
    m_struct_S, m_struct_1 = 2280, 8830
    T_eng_S, T_eng_1, T_eng_2 = 628500, 1083100, 110094
    I_S, I_1, I_2 = 283.33364, 301.68776, 467.21311
    tau_burn_S, tau_burn_1 = 75.2, 261
    m_payload = 4164
    m_t0_A = (9 * m_tot_S) + m_tot_1 + m_tot_2 + m_payload
    m_tF_A = m_t0_A - (6 * m_prop_S) - ((tau_burn_S / tau_burn_1) * m_prop_1)
    m_t0_B = m_tF_A - (6 * m_struct_S)
    m_tF_B = m_t0_B - (3 * m_prop_S) - ((tau_burn_S / tau_burn_1) * m_prop_1)
    m_t0_C = m_tF_B - (3 * m_struct_S)
    m_tF_C = m_t0_C - ((1 - (2 * (tau_burn_S / tau_burn_1))) * m_prop_1)
    m_t0_D = m_tF_C - m_struct_1
    m_tF_D = m_payload
    problem = _ocp_class(api)(name="Delta III Launch Vehicle Ascent Problem")
    A_ = -mu / (r_vec_norm ** 3)
    v_y_t0 = omega_E * R_E * sym.cos(psi_L)
    masses = [(m_t0_A, m_tF_A), (m_t0_B, m_tF_B), (m_t0_C, m_tF_C), (m_t0_D, m_tF_D)]
    times = [(t_launch, t_sep_S), (t_sep_S, t_sep_1), (t_sep_1, t_sep_2),
             (t_sep_2, t_orbit)]
    aux = [
        {T: (6 * T_eng_S) + T_eng_1,
         xi: (1 / g_0) * (6 * (T_eng_S / I_S) + (T_eng_1 / I_1))},
        {T: (3 * T_eng_S) + T_eng_1,
         xi: (1 / g_0) * ((3 * (T_eng_S / I_S)) + (T_eng_1 / I_1))},
        {T: T_eng_1, xi: T_eng_1 / (g_0 * I_1)},
        {T: T_eng_2, xi: T_eng_2 / (g_0 * I_2)}]
    phases = []
    for i, name in enumerate("ABCD"):
        ph = problem.new_phase(name)
        ph.state_variables = [r_x, r_y, r_z, v_x, v_y, v_z, m]
        ph.control_variables = [u_x, u_y, u_z]
        ph.state_equations = {
            r_x: v_x, r_y: v_y, r_z: v_z,
            v_x: (A_ * r_x) + (T_over_m * u_x) + (D_x / m),
            v_y: (A_ * r_y) + (T_over_m * u_y) + (D_y / m),
            v_z: (A_ * r_z) + (T_over_m * u_z) + (D_z / m),
            m: -xi}
        ph.path_constraints = [u_vec_norm - 1, r_vec_norm - R_E]
        ph.auxiliary_data = aux[i]
        ph.bounds.initial_time, ph.bounds.final_time = times[i]
        ph.bounds.state_variables = {
            r_x: [-2 * R_E, 2 * R_E], r_y: [-2 * R_E, 2 * R_E],
            r_z: [-2 * R_E, 2 * R_E], v_x: [-10000, 10000], v_y: [-10000, 10000],
            v_z: [-10000, 10000], m: [masses[i][1], masses[i][0]]}
        ph.bounds.control_variables = {u_x: [-1.1, 1.1], u_y: [-1.1, 1.1],
                                       u_z: [-1.1, 1.1]}
        ph.bounds.path_constraints = [[0, 0], [0, "inf"]]
        if i == 0:
            ph.bounds.initial_state_constraints = {
                r_x: R_E * sym.cos(psi_L), r_y: 0, r_z: R_E * sym.sin(psi_L),
                v_x: 0, v_y: v_y_t0, v_z: 0, m: m_t0_A}
        else:
            ph.bounds.initial_state_constraints = {m: masses[i][0]}
        ph.bounds.final_state_constraints = {m: masses[i][1]}
        ph.guess.time = list(times[i])
        ph.guess.state_variables = [
            [R_E * sym.cos(psi_L)] * 2, [0, 0], [R_E * sym.sin(psi_L)] * 2, [0, 0],
            [v_y_t0, v_y_t0], [0, 0], [masses[i][0], masses[i][1]]]
        ph.guess.control_variables = [[0.9, 0.9], [0.05, 0.05], [0.45, 0.45]]
        phases.append(ph)
    pD = phases[3]
    problem.objective_function = -(sym.sqrt(
        pD.final_state_variables.r_x ** 2 + pD.final_state_variables.r_y ** 2
        + pD.final_state_variables.r_z ** 2) - R_E)
    cons = []
    for p1, p2 in zip(phases[:-1], phases[1:]):
        for name in ("r_x", "r_y", "r_z", "v_x", "v_y", "v_z"):
            cons.append(getattr(p1.final_state_variables, name)
                        - getattr(p2.initial_state_variables, name))
    problem.endpoint_constraints = cons
    problem.bounds.endpoint_constraints = [0] * len(cons)
    problem.auxiliary_data = {
        mu: 3.986012e14, R_E: 6378145,
        r_vec_norm: sym.sqrt(r_x ** 2 + r_y ** 2 + r_z ** 2),
        v_vec_norm: sym.sqrt(v_x ** 2 + v_y ** 2 + v_z ** 2),
        u_vec_norm: sym.sqrt(u_x ** 2 + u_y ** 2 + u_z ** 2),
        D_x: -0.5 * C_D * S * rho * v_r_vec_norm * v_r_x,
        D_y: -0.5 * C_D * S * rho * v_r_vec_norm * v_r_y,
        D_z: -0.5 * C_D * S * rho * v_r_vec_norm * v_r_z,
        C_D: 0.5, S: 4 * np.pi,
        v_r_vec_norm: sym.sqrt(v_r_x ** 2 + v_r_y ** 2 + v_r_z ** 2),
        v_r_x: v_x - oxr_x, v_r_y: v_y - oxr_y, v_r_z: v_z - oxr_z,
        oxr_x: -omega_E * r_y, oxr_y: omega_E * r_x, oxr_z: 0,
        g_0: 9.80665, h_0: 7200, h: r_vec_norm - R_E,
        rho: rho_0 * sym.exp(-h / h_0), rho_0: 1.225, omega_E: 7.29211585e-5,
        T_over_m: T / m, psi_L: (28.5 / 180) * np.pi}
    problem.settings.quadrature_method = "lobatto"
    return problem

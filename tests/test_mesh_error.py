"""Mesh-refinement error.

Golden fixtures ``tests/golden/mesh_error_*.npz`` hold the outputs of the
reference's own ``PattersonRaoMeshRefinement.phase_mesh_error``
(``pycollo/mesh_refinement.py:206-240``) executed on the reference's own ph mesh
(``oracle/make_golden.py``).  CPU: the oracle restatement reproduces them
bit-for-bit.  GPU: ``pcx_mesh_error`` reproduces them to 1e-12 relative /
1e-14 absolute from ``x_ph`` alone (it evaluates dy_ph itself)."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import GOLDEN
from oracle import mesh_error as OM
from examples import problems as examples
from pycollo_b200.mesh import Mesh, PhaseMesh
from pycollo_b200.quadrature import Quadrature

CASES = [("robot_lobatto", "free_flying_robot", "lobatto"),
         ("robot_radau", "free_flying_robot", "radau"),
         ("shuttle_lobatto", "space_shuttle_reentry", "lobatto"),
         ("multiphase_lobatto", "multiphase_sliding_mass", "lobatto")]


def _load(tag):
    return np.load(f"{GOLDEN}/mesh_error_{tag}.npz")


def _problem(name, method):
    ocp = getattr(examples, name)()
    ocp.settings.quadrature_method = method
    ocp.settings.scaling_method = "none"
    return ocp


@pytest.mark.parametrize("tag,problem,method", CASES)
def test_oracle_reproduces_reference_mesh_error(tag, problem, method):
    g = _load(tag)
    ocp = _problem(problem, method)
    nodes_ph = OM.ph_section_nodes(g["section_nodes"])
    quad = Quadrature.adopt(np.load(f"{GOLDEN}/quadrature_{method}.npz"), method)
    ph = Mesh(quad, [PhaseMesh(len(nodes_ph), g["section_sizes"], nodes_ph)
                     for _ in ocp.phases], 2, 17)
    from pycollo_b200.backend import lower_problem
    low = lower_problem(ocp, ph.p)
    for ip, (irp, t) in enumerate(zip(low.ir.phases, low.S.ph)):
        N = t.N
        sI = sp.csr_matrix((g[f"sI_data_{ip}"], g[f"sI_indices_{ip}"], g[f"sI_indptr_{ip}"]),
                           shape=(N - 1, N))
        # with the reference's quadrature tables adopted, the ph mesh operators
        # ARE the reference's (same single rounded product A * h_k, mesh.py:300)
        assert np.max(np.abs((ph.sI_matrix[ip] - sI).toarray())) <= 1e-16
        own = Mesh(Quadrature(method), [PhaseMesh(len(nodes_ph), g["section_sizes"], nodes_ph)], 2, 17)
        assert np.max(np.abs((own.sI_matrix[0] - sI).toarray())) < 3e-12   # own tables: see DESIGN.md section 1
        y_ph = g["x_ph"][t.x_off:t.x_off + irp.n_y * N].reshape(irp.n_y, N)
        a, r, m = OM.phase_mesh_error(g[f"dy_{ip}"], y_ph, sI, float(g[f"stretch_{ip}"]),
                                      nodes_ph, ph.mesh_index_boundaries[ip])
        assert np.array_equal(a, g[f"abs_{ip}"])
        assert np.array_equal(r, g[f"rel_{ip}"])
        assert np.array_equal(m, g[f"max_{ip}"])
        assert m.max() > 1e-6          # the fixture is not trivially zero


def test_ph_mesh_has_one_more_node_per_section():
    from pycollo_b200.mesh_refinement import create_ph_mesh
    quad = Quadrature("lobatto")
    mesh = Mesh(quad, [PhaseMesh(4, [0.1, 0.2, 0.3, 0.4], [3, 5, 2, 4])])
    ph = create_ph_mesh(mesh)
    assert list(ph.N_K[0]) == [4, 6, 3, 5]
    assert np.allclose(ph.h_K[0], mesh.h_K[0])
    assert ph.N[0] == mesh.N[0] + 4
    # section boundaries coincide (mesh_refinement.py:164-166)
    assert np.allclose(ph.tau[0][ph.mesh_index_boundaries[0]],
                       mesh.tau[0][mesh.mesh_index_boundaries[0]])


def test_polynomial_refit_reproduces_polynomial_solutions():
    """solution_abc.py:60-101: on a section of N_k nodes a degree N_k-1 state
    derivative is re-fitted exactly, so y_ph equals the true polynomial."""
    quad = Quadrature("lobatto")
    mesh = Mesh(quad, [PhaseMesh(3, [0.2, 0.5, 0.3], [4, 5, 3])])
    tau = mesh.tau[0]
    T = 3.0
    coef = np.array([0.3, -1.0, 0.5])                     # y = c0 + c1 t + c2 t^2, t = tau*T/2
    t = tau * T / 2
    y = (coef[0] + coef[1] * t + coef[2] * t * t)[None, :]
    dy = (coef[1] + 2 * coef[2] * t)[None, :]
    u = np.cos(tau)[None, :]
    yp, up = OM.fit_section_polys(tau, y, dy, u, T, mesh.mesh_index_boundaries[0], mesh.N_K[0])
    from pycollo_b200.mesh_refinement import create_ph_mesh
    ph = create_ph_mesh(mesh)
    y_ph = OM.interpolate_to_ph(y, yp, mesh.mesh_index_boundaries[0],
                                ph.mesh_index_boundaries[0], ph.tau[0])
    t_ph = ph.tau[0] * T / 2
    assert np.allclose(y_ph[0], coef[0] + coef[1] * t_ph + coef[2] * t_ph ** 2, atol=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,problem,method", CASES)
def test_gpu_mesh_error_matches_reference(tag, problem, method):
    from pycollo_b200.mesh_refinement import MeshErrorEvaluator
    g = _load(tag)
    ocp = _problem(problem, method)
    quad = Quadrature.adopt(np.load(f"{GOLDEN}/quadrature_{method}.npz"), method)
    mesh = Mesh(quad, [PhaseMesh(len(g["section_nodes"]), g["section_sizes"], g["section_nodes"])
                       for _ in ocp.phases], 2, 16)
    ev = MeshErrorEvaluator(ocp, mesh)
    res = ev(g["x_ph"])
    for ip, (a, r, m) in enumerate(res):
        ea, er, em = g[f"abs_{ip}"], g[f"rel_{ip}"], g[f"max_{ip}"]
        assert a.shape == ea.shape and m.shape == em.shape
        # 1e-12 relative / 1e-14 absolute (BASELINE north star), scaled by |Y|
        yscale = 1.0 + np.max(np.abs(g["x_ph"]))
        assert np.max(np.abs(a - ea)) <= 1e-12 * yscale
        assert np.max(np.abs(r - er)) <= 1e-12
        assert np.max(np.abs(m - em)) <= 1e-12
        assert np.all(a[ea == 0.0] == 0.0)             # unused tail stays zero


@pytest.mark.gpu
def test_gpu_mesh_refinement_from_solution():
    """Solution -> construct_x_ph -> device error pass against the oracle chain."""
    from oracle.blockwise import BlockwiseNLP
    from pycollo_b200.backend import Cuda
    from pycollo_b200.solution import Solution
    ocp = examples.free_flying_robot()
    ocp.settings.scaling_method = "none"
    examples.set_mesh(ocp, 12, 5)
    backend = Cuda(ocp)
    for step in ("create_bounds", "create_scaling", "create_quadrature",
                 "create_initial_mesh", "create_guess", "create_mesh_iterations"):
        getattr(backend, step)()
    it = backend.current_iteration
    it.generate_nlp()
    rng = np.random.default_rng(5)
    x = it.guess_x_tilde + 0.05 * rng.standard_normal(it.S.num_x)
    sol = Solution(it, x)
    mr = sol.refine_mesh()
    # oracle chain on the same x
    ip = 0
    pd = sol.phase_data[ip]
    mesh = it.mesh
    yp, up = OM.fit_section_polys(pd.tau, pd.y, pd.dy, pd.u, pd.T,
                                  mesh.mesh_index_boundaries[ip], mesh.N_K[ip])
    ph = mr.ph_mesh
    y_ph = OM.interpolate_to_ph(pd.y, yp, mesh.mesh_index_boundaries[ip],
                                ph.mesh_index_boundaries[ip], ph.tau[ip])
    u_ph = OM.interpolate_to_ph(pd.u, up, mesh.mesh_index_boundaries[ip],
                                ph.mesh_index_boundaries[ip], ph.tau[ip])
    # device re-fit (pcx_refit_to_ph) vs the reference-style numpy fits: order 5
    scale = 1.0 + np.abs(pd.y).max()
    assert np.max(np.abs(y_ph - mr.y_ph[ip])) <= 1e-12 * scale
    assert np.max(np.abs(u_ph - mr.u_ph[ip])) <= 1e-12 * (1.0 + np.abs(pd.u).max())
    # ... and vs the host mirror that uses the same numpy polynomial objects
    x_ph_host, y_host, u_host = mr.construct_x_ph()
    assert x_ph_host.shape == mr.x_ph.shape
    assert np.max(np.abs(x_ph_host - mr.x_ph)) <= 1e-12 * (1.0 + np.abs(x_ph_host).max())
    low = mr.evaluator.low
    B = BlockwiseNLP(ocp, low.ir.full_bounds,
                     [dict(N=m.N, sI=m.sI_matrix, sA=m.sA_matrix, W=m.W_matrix) for m in ph.p],
                     scaling_method="none")
    dy = B.dy(mr.x_ph)
    a, r, m = OM.phase_mesh_error(dy[:pd.y.shape[0] * ph.N[ip]], mr.y_ph[ip], ph.sI_matrix[ip],
                                  pd.stretch, ph.N_K[ip], ph.mesh_index_boundaries[ip])
    assert np.max(np.abs(mr.absolute_mesh_errors[ip] - a)) <= 1e-12 * (1 + np.abs(mr.x_ph).max())
    assert np.max(np.abs(mr.maximum_relative_mesh_errors[ip] - m)) <= 1e-12


REFIT_CASES = [("free_flying_robot", "lobatto", [4, 6, 3, 8, 5, 4, 9, 2]),
               ("free_flying_robot", "radau", [4, 6, 3, 8, 5, 4, 9, 3]),
               ("multiphase_sliding_mass", "lobatto", [3, 5, 4, 6, 2, 7, 4, 5]),
               ("space_shuttle_reentry", "radau", [5, 4, 7, 3, 6, 4, 3, 8])]


def test_refit_matrices_are_exact_on_polynomials():
    """quadrature.refit_matrices: integral of the interpolant / interpolant at the
    interior ph nodes, exact (1e-14) for polynomial data of the section's degree,
    where the reference's numpy fits in the [0,1] window lose digits with the order."""
    for method in ("lobatto", "radau"):
        q = Quadrature(method)
        for n in (2, 3, 4, 7, 10, 16):
            Cy, Pu = q.refit_matrices(n)
            xi = np.array(q.quadrature_point(n), dtype=float)
            zeta = np.array(q.quadrature_point(n + 1), dtype=float)
            if method == "radau":
                xi[-1] = zeta[-1] = 1.0
            zin = zeta[1:n]
            deg_y = n - 1 if method == "lobatto" else n - 2
            for d in range(deg_y + 1):
                exact = (zin ** (d + 1) - (-1.0) ** (d + 1)) / (d + 1)
                assert np.max(np.abs(Cy @ xi ** d - exact)) <= 2e-14, (method, n, d)
            for d in range(n):
                assert np.max(np.abs(Pu @ xi ** d - zin ** d)) <= 2e-14, (method, n, d)
            if method == "radau":
                assert np.all(Cy[:, -1] == 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("problem,method,nodes", REFIT_CASES)
def test_gpu_refit_to_ph_matches_reference_fits(problem, method, nodes):
    """pcx_refit_to_ph vs the oracle's restatement of the reference's numpy fits
    (solution_abc.py:60-142, mesh_refinement.py:160-204).  Tolerance: the numpy
    fits themselves are only good to ~1e-12 at order 9-10 (test above)."""
    from pycollo_b200 import engine as E
    from pycollo_b200.backend import lower_problem
    from pycollo_b200.mesh_refinement import create_ph_mesh
    ocp = _problem(problem, method)
    sizes = [0.1, 0.15, 0.05, 0.2, 0.1, 0.12, 0.08, 0.2]
    mesh = Mesh(Quadrature(method), [PhaseMesh(len(nodes), sizes, nodes) for _ in ocp.phases], 2, 16)
    low = lower_problem(ocp, mesh.p)
    S = low.S
    eng = E.Engine(S, low.layouts, low.header)
    eng.set_scaling(np.ones(S.n_var_ocp), np.zeros(S.n_var_ocp), np.ones(S.n_con_ocp), 1.0)
    rng = np.random.default_rng(8)
    x = rng.uniform(0.2, 0.6, S.num_x)
    dy = eng.eval_host(E.EVAL_DY, x)["dy"][0]
    x_ph = eng.refit_to_ph_host(x, dy)
    ph = create_ph_mesh(mesh, 2, 16)
    off = 0
    for ip, (irp, t) in enumerate(zip(low.ir.phases, S.ph)):
        N, Nph = t.N, ph.N[ip]
        ny, nu = irp.n_y, irp.n_u
        y = x[t.x_off:t.x_off + ny * N].reshape(ny, N)
        u = x[t.x_off + ny * N:t.x_off + (ny + nu) * N].reshape(nu, N)
        d = dy[t.dy_off:t.dy_off + ny * N].reshape(ny, N)
        tv = x[t.q_col + irp.n_q:t.q_col + irp.n_q + irp.n_t]
        t0 = tv[0] if irp.t_needed[0] else float(irp.t0)
        tF = tv[-1] if irp.t_needed[1] else float(irp.tF)
        bnd, bph = mesh.mesh_index_boundaries[ip], ph.mesh_index_boundaries[ip]
        yp, up = OM.fit_section_polys(mesh.tau[ip], y, d, u, tF - t0, bnd, mesh.N_K[ip], method)
        y_ref = OM.interpolate_to_ph(y, yp, bnd, bph, ph.tau[ip])
        u_ref = OM.interpolate_to_ph(u, up, bnd, bph, ph.tau[ip])
        got_y = x_ph[off:off + ny * Nph].reshape(ny, Nph)
        got_u = x_ph[off + ny * Nph:off + (ny + nu) * Nph].reshape(nu, Nph)
        tol = 5e-12 * (1.0 + max(np.abs(y_ref).max(), np.abs(d).max() * abs(tF - t0)))
        assert np.max(np.abs(got_y - y_ref)) <= tol
        assert np.max(np.abs(got_u - u_ref)) <= 5e-12 * (1.0 + np.abs(u_ref).max())
        assert np.array_equal(got_y[:, bph], y[:, bnd])        # boundary values are copied
        nqt = irp.n_q + irp.n_t
        assert np.array_equal(x_ph[off + (ny + nu) * Nph:off + (ny + nu) * Nph + nqt],
                              x[t.q_col:t.q_col + nqt])
        off += (ny + nu) * Nph + nqt
    assert np.array_equal(x_ph[off:], x[S.s_off:])
    assert off + S.NS == eng.refit_size()


@pytest.mark.gpu
@pytest.mark.parametrize("problem,K,nodes", [("free_flying_robot", 40, 4),
                                             ("multiphase_sliding_mass", 24, [3, 5, 4] * 8)])
def test_gpu_mesh_error_sharded_by_section(problem, K, nodes):
    """the error pass of one mesh split over three
    "ranks" (three engines restricted to disjoint tile ranges on one device).  It is
    section-local -- no exchange: every section is evaluated by exactly one rank,
    the union of the ranks' entries is the unsharded result bit for bit, and the
    global maximum is the maximum of the local maxima."""
    from pycollo_b200.mesh import Mesh, PhaseMesh
    from pycollo_b200.mesh_refinement import MeshErrorEvaluator
    from pycollo_b200.quadrature import Quadrature
    ocp = getattr(examples, problem)()
    ocp.settings.scaling_method = "none"
    mesh = Mesh(Quadrature("lobatto"), [PhaseMesh(K, None, nodes) for _ in ocp.phases], 2, 20)
    whole = MeshErrorEvaluator(ocp, mesh, max_tile_nodes=24)
    rng = np.random.default_rng(8)
    x_ph = rng.uniform(0.1, 0.9, whole.low.S.num_x)
    ref = whole(x_ph)
    world = 3
    parts = [MeshErrorEvaluator(ocp, mesh, shard=(r, world), max_tile_nodes=24) for r in range(world)]
    assert whole.low.S.num_tiles >= world
    got = [p(x_ph) for p in parts]
    for ip, (a_ref, r_ref, m_ref) in enumerate(ref):
        owner = np.zeros(len(m_ref), dtype=int)
        a_sum, r_sum, m_sum = np.zeros_like(a_ref), np.zeros_like(r_ref), np.zeros_like(m_ref)
        for p, res in zip(parts, got):
            lo, hi = p.local_sections[ip]
            owner[lo:hi] += 1
            a, r, m = res[ip]
            assert not a[:lo].any() and not a[hi:].any() and not m[:lo].any() and not m[hi:].any()
            a_sum += a; r_sum += r; m_sum += m
        assert np.all(owner == 1)                       # one rank per section
        np.testing.assert_array_equal(a_sum, a_ref)
        np.testing.assert_array_equal(r_sum, r_ref)
        np.testing.assert_array_equal(m_sum, m_ref)
    worst = max(p.global_maximum([res[ip][2] for ip in range(len(ref))]) for p, res in zip(parts, got))
    assert worst == max(float(m.max()) for _, _, m in ref)


SOLUTION_CASES = ["brachistochrone_lobatto", "brachistochrone_lobatto_ragged", "cart_pole_radau",
                  "free_flying_robot_lobatto", "multiphase_lobatto", "hypersensitive_radau"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", SOLUTION_CASES)
def test_gpu_solution_chain_matches_executed_reference(name):
    """Rows a11 / N2 / a12 against the reference ITSELF (``oracle/make_golden_solution.py``
    executes the unmodified ``CasadiSolution`` -> ``interpolate_solution_*`` ->
    ``PattersonRaoMeshRefinement`` chain on a synthetic iterate): state derivatives
    (``PCX_EVAL_DY``), the solution re-fitted onto the p+1 mesh on the device
    (``pcx_refit_to_ph``) and the mesh errors (``pcx_mesh_error``)."""
    from examples.cases import build_golden_problem
    from helpers import max_err, strict_err
    from pycollo_b200.solution import Solution
    g = np.load(f"{GOLDEN}/solution_{name}.npz")
    ocp = build_golden_problem(name)
    ocp.initialise()
    it = ocp._backend.mesh_iterations[0]
    sol = Solution(it, g["x"], J=None)
    dy = np.concatenate([np.ravel(p.dy) for p in sol.phase_data])
    e_dy = max_err(dy, g["dy"])
    mr = sol.refine_mesh()
    e_xph = max_err(mr.x_ph, g["x_ph"])
    worst = 0.0
    for ip in range(int(g["num_phases"])):
        np.testing.assert_array_equal(mr.ph_mesh.tau[ip], g[f"tau_ph_{ip}"])
        for ours, key in ((mr.absolute_mesh_errors[ip], "abs"), (mr.relative_mesh_errors[ip], "rel"),
                          (mr.maximum_relative_mesh_errors[ip], "max")):
            worst = max(worst, max_err(ours, g[f"{key}_{ip}"]))
    print(f"solution chain {name}: dy {e_dy:.1e}  x_ph {e_xph:.1e}  mesh errors {worst:.1e}")
    # the reference fits least-squares polynomials in a [0, 1] window (loses ~1e-13 at
    # order 6); the device applies the exact interpolation matrices
    assert e_dy <= 1e-12 and e_xph <= 1e-11 and worst <= 1e-9

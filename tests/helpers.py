"""Shared builders for the tests (problem + mesh -> lowered problem, oracle, engine)."""
import numpy as np
import scipy.sparse as sp

from pycollo_b200 import engine as E
from pycollo_b200.backend import lower_problem
from pycollo_b200.mesh import PhaseMesh, PhaseMeshData
from pycollo_b200.quadrature import Quadrature

RAGGED_SIZES = [0.05, 0.2, 0.1, 0.15, 0.3, 0.2]
RAGGED_NODES = [4, 7, 2, 10, 3, 5]
GOLDEN = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden")


def make_meshes(ocp, method, K, nodes, sizes=None):
    ocp.settings.quadrature_method = method
    quad = Quadrature(method)
    return [PhaseMeshData(quad, PhaseMesh(K, sizes, nodes), 2, 20) for _ in ocp.phases]


def oracle_meshes(meshes):
    return [dict(N=m.N, sI=m.sI_matrix, sA=m.sA_matrix, W=m.W_matrix) for m in meshes]


def golden_mesh(name):
    m = np.load(f"{GOLDEN}/{name}.npz")
    N = int(m["N"])
    sI = sp.csr_matrix((m["sI_data"], m["sI_indices"], m["sI_indptr"]), shape=(N - 1, N))
    sA = sp.csr_matrix((m["sA_data"], m["sA_indices"], m["sA_indptr"]), shape=(N - 1, N))
    return dict(N=N, sI=sI, sA=sA, W=m["W"], tau=m["tau"], h_K=m["h_K"],
                N_K=m["section_nodes"], sizes=m["section_sizes"])


def build_case(ocp, method, K, nodes, sizes=None, seed=0, unit_scaling=False,
               oracle=True, **structure_kwargs):
    """Lowered problem + oracle with random constraint/objective scaling."""
    from examples.cases import lower_case
    low, meshes, scal = lower_case(ocp, method, K, nodes, sizes, seed, unit_scaling,
                                   **structure_kwargs)
    B = None
    if oracle:
        from oracle.blockwise import BlockwiseNLP
        B = BlockwiseNLP(ocp, low.ir.full_bounds, oracle_meshes(meshes), W_ocp=scal[2],
                         w=scal[3], prune=low.S.prune,
                         scaling_method=ocp.settings.scaling_method)
        np.testing.assert_array_equal(B.V, scal[0])
        np.testing.assert_array_equal(B.r, scal[1])
    return low, B, scal


def make_engine(low, scal, batch=1):
    eng = E.Engine(low.S, low.layouts, low.header, batch=batch)
    eng.set_scaling(scal[0], scal[1], scal[2], scal[3])
    return eng


def strict_err(a, b, b_err=None):
    """Worst deviation of ``a`` from the reference values ``b`` in units where
    ``<= 1e-12`` means exactly north_star's bar, element by element:
    ``|a-b| <= 1e-12*|b|``  OR  ``|a-b| <= 1e-14``.

    ``b_err`` (golden files only): per element, the first-order running-error bound
    of evaluating the REFERENCE's own expression in fp64 (unit roundoff per
    operation, exact inputs; ``oracle/refshim/casadi.error_bounds``).  Where that
    bound exceeds the bar -- cancellation: ``x = V*x_tilde + r`` loses ``eps*|r|``
    before anything else happens -- the reference's own evaluation is not defined
    more precisely than it, and the element is measured against 4x the bound."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    assert np.all(np.isfinite(a)), f"{np.sum(~np.isfinite(a))} non-finite values"
    d = np.abs(a - b)
    with np.errstate(divide="ignore", invalid="ignore"):
        units = np.minimum(np.where(b != 0, d / np.abs(b), np.inf), d * 100.0)
    if b_err is not None:
        noise = 4.0 * np.asarray(b_err, dtype=float).reshape(b.shape)
        units = np.where(d <= noise, np.minimum(units, 1e-12), units)
    return float(np.max(units))


def max_err(a, b):
    """Comparison of two fp64 IMPLEMENTATIONS (CUDA against the numpy oracle, sharded
    against unsharded ...), where neither side is exact and both carry the rounding
    noise of their own operation order: relative error where the value is
    significant, absolute in units of 1 % of the vector's own scale where
    cancellation leaves only that noise.  The strict north_star bar (``strict_err``)
    is applied where an exactly rounded reference exists: the golden files."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    assert np.all(np.isfinite(a)), f"{np.sum(~np.isfinite(a))} non-finite values"
    scale = max(1.0, float(np.max(np.abs(b)))) * 1e-2
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), scale)))


def cpu_mesh_error_chain(ocp, low, low_ph, mesh, ph, xh, dyh):
    """CPU chain of the reference's mesh-error algorithm (oracle/mesh_error.py:
    numpy polynomial fits per state and section, then the error loops) for a
    solution ``xh`` / ``dyh`` on ``mesh``; returns the largest relative error.
    Used by tests and by tools/mesh_error_bench.py as the timed CPU leg."""
    from oracle import mesh_error as OM
    from oracle.blockwise import BlockwiseNLP
    S = low.S
    B = BlockwiseNLP(ocp, low_ph.ir.full_bounds,
                     [dict(N=m.N, sI=m.sI_matrix, sA=m.sA_matrix, W=m.W_matrix) for m in ph.p],
                     scaling_method="none")
    sI = [m.sI_matrix for m in ph.p]
    xs, Ts = [], []
    for ip, (irp, t) in enumerate(zip(low.ir.phases, S.ph)):
        ny, nu, N = irp.n_y, irp.n_u, t.N
        y = xh[t.x_off:t.x_off + ny * N].reshape(ny, N)
        u = xh[t.x_off + ny * N:t.x_off + (ny + nu) * N].reshape(nu, N)
        d = dyh[t.dy_off:t.dy_off + ny * N].reshape(ny, N)
        tv = xh[t.q_col + irp.n_q:t.q_col + irp.n_q + irp.n_t]
        T = (tv[-1] if irp.t_needed[1] else float(irp.tF)) - (tv[0] if irp.t_needed[0] else float(irp.t0))
        Ts.append(T)
        bnd, bph = mesh.mesh_index_boundaries[ip], ph.mesh_index_boundaries[ip]
        yp, up = OM.fit_section_polys(mesh.tau[ip], y, d, u, T, bnd, mesh.N_K[ip])
        y_ph = OM.interpolate_to_ph(y, yp, bnd, bph, ph.tau[ip])
        u_ph = OM.interpolate_to_ph(u, up, bnd, bph, ph.tau[ip])
        xs += [y_ph.ravel(), u_ph.ravel(), xh[t.q_col:t.q_col + irp.n_q + irp.n_t]]
    xs.append(xh[S.s_off:])
    x_ph = np.concatenate(xs)
    dyp = B.dy(x_ph)
    o, worst = 0, 0.0
    for ip, (irp, t) in enumerate(zip(low.ir.phases, low_ph.S.ph)):
        ny, Nph = irp.n_y, t.N
        y_ph = x_ph[t.x_off:t.x_off + ny * Nph].reshape(ny, Nph)
        _, _, m = OM.phase_mesh_error(dyp[o:o + ny * Nph], y_ph, sI[ip], 0.5 * Ts[ip], ph.N_K[ip],
                                      ph.mesh_index_boundaries[ip])
        o += ny * Nph
        worst = max(worst, float(m.max()))
    return worst

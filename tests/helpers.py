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
    return [PhaseMeshData(quad, PhaseMesh(K, sizes, nodes), 2, 16) for _ in ocp.phases]


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
    from oracle.blockwise import BlockwiseNLP
    meshes = make_meshes(ocp, method, K, nodes, sizes)
    low = lower_problem(ocp, meshes, **structure_kwargs)
    rng = np.random.default_rng(seed)
    W_ocp = np.ones(low.S.n_con_ocp) if unit_scaling else \
        rng.uniform(0.5, 2.0, low.S.n_con_ocp)
    w = 1.0 if unit_scaling else 1.7
    B = None
    if oracle:
        B = BlockwiseNLP(ocp, low.ir.full_bounds, oracle_meshes(meshes), W_ocp=W_ocp,
                         w=w, prune=low.S.prune,
                         scaling_method=ocp.settings.scaling_method)
        V, r = B.V, B.r
    else:
        from pycollo_b200.backend import Bounds
        bnd = Bounds(low.ir)
        V = bnd.x_bnd_upper - bnd.x_bnd_lower
        r = bnd.x_bnd_upper - V / 2
    return low, B, (V, r, W_ocp, w)


def make_engine(low, scal, batch=1):
    eng = E.Engine(low.S, low.layouts, low.header, batch=batch)
    eng.set_scaling(scal[0], scal[1], scal[2], scal[3])
    return eng


def max_err(a, b):
    """Worst |a-b| measured against max(1e-2*|b|_inf-scale, |b|): relative error
    where the value is significant, absolute (in units of the vector's own scale)
    where cancellation leaves only rounding noise."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    assert np.all(np.isfinite(a)), f"{np.sum(~np.isfinite(a))} non-finite values"
    scale = max(1.0, float(np.max(np.abs(b)))) * 1e-2
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), scale)))

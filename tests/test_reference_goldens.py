"""Parity against the REFERENCE ITSELF: golden vectors emitted by executing the
unmodified ``/root/reference/pycollo`` package (``oracle/make_golden_nlp.py``:
``Casadi.generate_nlp_function_callables`` ``backend.py:1403-1679``,
``evaluate_G_structure`` ``:1747-1761``, the nlpsol Hessian ``:1693``,
``IterationScaling`` ``scaling.py:124-454``, guess interpolation and bound
expansion ``iteration.py:86-194, 396-453``) on the same user scripts
(``examples/problems.py``) built on the reference's own API.

* sparsity indices: bit-exact (G in CCS order, H upper triangle in CCS order);
* values: ``|a-b| <= 1e-12*|b|`` or ``<= 1e-14`` element by element
  (``helpers.strict_err``; an element whose own fp64 evaluation by the reference is
  ill-conditioned is held to 4x the reference's running-error bound instead),
  observed maxima printed;
* CPU tests pin the oracle and the host-side mirror (mesh tables bit for bit,
  V/r, scaled guess, bounds); ``-m gpu`` tests pin the CUDA path through the
  ``Cuda`` backend object / C ABI, including the scaling it derives on the device.
"""
import os

import numpy as np
import pytest

from examples.cases import GOLDEN_CASES, build_golden_problem
from helpers import GOLDEN, strict_err
from pycollo_b200.backend import lower_problem
from pycollo_b200.mesh import PhaseMeshData
from pycollo_b200.quadrature import Quadrature

RTOL = 1e-12
NAMES = [n for n in GOLDEN_CASES if os.path.exists(f"{GOLDEN}/nlp_{n}.npz")]
SMALL = [n for n in NAMES if n not in ("double_pendulum_lobatto", "delta_iii_lobatto",
                                       "shuttle_lobatto")]


def golden(name):
    return np.load(f"{GOLDEN}/nlp_{name}.npz")


def our_meshes(ocp):
    quad = Quadrature(ocp.settings.quadrature_method)
    return [PhaseMeshData(quad, ph.mesh, ocp.settings.collocation_points_min, 20)
            for ph in ocp.phases]


def test_every_case_has_a_golden_file():
    assert set(NAMES) == set(GOLDEN_CASES), sorted(set(GOLDEN_CASES) - set(NAMES))


@pytest.mark.parametrize("name", NAMES)
def test_mesh_tables_bit_identical(name):
    """tau, W and the integration / difference CSRs of ``pycollo/mesh.py:236-356``
    built from ``quadrature.py:116-261``: equal to the last bit (default tables)."""
    g = golden(name)
    ocp = build_golden_problem(name)
    for ip, m in enumerate(our_meshes(ocp)):
        np.testing.assert_array_equal(m.tau, g[f"tau_{ip}"])
        np.testing.assert_array_equal(m.N_K, g[f"N_K_{ip}"])
        np.testing.assert_array_equal(m.h_K, g[f"h_K_{ip}"])
        np.testing.assert_array_equal(m.W_matrix, g[f"Wq_{ip}"])
        for ours, tag in ((m.sI_matrix, "sI"), (m.sA_matrix, "sA")):
            ours = ours.tocsr()
            np.testing.assert_array_equal(ours.indptr, g[f"{tag}_{ip}_indptr"])
            np.testing.assert_array_equal(ours.indices, g[f"{tag}_{ip}_indices"])
            np.testing.assert_array_equal(ours.data, g[f"{tag}_{ip}_data"])


@pytest.mark.parametrize("name", NAMES)
def test_structure_bit_exact(name):
    """Sizes and sparsity of G and H: the engine's device-table builder
    (``structure.py``) against the reference's own patterns."""
    g = golden(name)
    ocp = build_golden_problem(name)
    S = lower_problem(ocp, our_meshes(ocp)).S
    assert (S.num_x, S.num_c) == (int(g["num_x"]), int(g["num_c"]))
    assert (S.nnz_g, S.nnz_h) == (int(g["G_nnz"]), int(g["H_nnz"]))
    rows, cols = S.G_structure()
    np.testing.assert_array_equal(rows, g["G_row"])
    np.testing.assert_array_equal(cols, g["G_col"])
    rows, cols = S.H_structure()
    np.testing.assert_array_equal(rows, g["H_row"])
    np.testing.assert_array_equal(cols, g["H_col"])


def assert_guess_equal(name, ours, ref):
    """Bit-exact, except where the user's guess is SYMBOLIC (Delta III:
    ``R_E*cos(psi_L)``): the stand-in evaluates ``cos`` with mpmath (correctly
    rounded) while libm -- CasADi's and this engine's -- may differ by one ulp."""
    if name.startswith("delta_iii"):
        np.testing.assert_allclose(ours, ref, rtol=4.5e-16, atol=1.2e-16)
    else:
        np.testing.assert_array_equal(ours, ref)


def _oracle(name, g, cls=None):
    from oracle.blockwise import BlockwiseNLP
    ocp = build_golden_problem(name)
    meshes = our_meshes(ocp)
    low = lower_problem(ocp, meshes)
    B = (cls or BlockwiseNLP)(
        ocp, low.ir.full_bounds,
        [dict(N=m.N, sI=m.sI_matrix, sA=m.sA_matrix, W=m.W_matrix) for m in meshes],
        W_ocp=g["W_ocp"], w=float(g["w"]), prune=low.S.prune,
        scaling_method=ocp.settings.scaling_method)
    return ocp, low, B


def _check_values(name, g, fns, label):
    worst = {}
    for k in range(len(g["x"])):
        x, lam, sg = g["x"][k], g["lam"][k], float(g["sigma"][k])
        got = fns(x, lam, sg)
        for key, val in got.items():
            ref = g[key][k]
            e = strict_err(np.reshape(val, np.shape(ref)), ref, g[key + "_err"][k])
            worst[key] = max(worst.get(key, 0.0), e)
    print(f"{label} {name}: " + "  ".join(f"{k} {v:.1e}" for k, v in worst.items()))
    bad = {k: v for k, v in worst.items() if v > RTOL}
    assert not bad, bad


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference(name):
    """``oracle/blockwise.py`` (the checker of every other CUDA test and the timed
    CPU baseline): patterns, V/r and all six callbacks against the reference."""
    g = golden(name)
    ocp, low, B = _oracle(name, g)
    np.testing.assert_array_equal(B.V, g["V_ocp"])
    np.testing.assert_array_equal(B.r, g["r_ocp"])
    rows, cols = B.G_structure()
    np.testing.assert_array_equal(rows, g["G_row"])
    np.testing.assert_array_equal(cols, g["G_col"])
    rows, cols = B.H_structure()
    np.testing.assert_array_equal(rows, g["H_row"])
    np.testing.assert_array_equal(cols, g["H_col"])
    _check_values(name, g, lambda x, lam, sg: dict(
        J=B.J(x), g=B.g(x), c=B.c(x), dy=B.dy(x), G_data=B.G_nonzeros(x),
        H_data=B.H_nonzeros(x, sg, lam)), "oracle")


@pytest.mark.parametrize("name", ["brachistochrone_radau", "hypersensitive_radau",
                                  "brachistochrone_lobatto_ragged"])
def test_expanded_oracle_matches_reference(name):
    """``oracle/expand.py`` (the literal statement-by-statement restatement) too."""
    from oracle.expand import ExpandedNLP
    g = golden(name)
    ocp, low, E = _oracle(name, g, ExpandedNLP)
    rows, cols = E.G_structure()
    np.testing.assert_array_equal(rows, g["G_row"])
    np.testing.assert_array_equal(cols, g["G_col"])
    rows, cols = E.H_structure()
    np.testing.assert_array_equal(rows, g["H_row"])
    np.testing.assert_array_equal(cols, g["H_col"])
    _check_values(name, g, lambda x, lam, sg: dict(
        J=E.J(x), c=E.c(x), G_data=E.G_nonzeros(x), H_data=E.H_nonzeros(x, sg, lam)),
        "expand")


@pytest.mark.parametrize("name", NAMES)
def test_host_iteration_matches_reference(name):
    """What the iteration derives before the first callback (rows a1, a2, N3):
    x layout and sizes, V/r on the mesh, the scaled initial guess
    (``iteration.py:86-194, 360-373``) and the scaled variable / constraint bounds
    (``:396-453``) -- host mirror, no device needed."""
    g = golden(name)
    ocp = build_golden_problem(name)
    ocp.settings.defer_engine = True
    ocp.initialise()
    it = ocp._backend.mesh_iterations[0]
    assert (it.num_x, it.num_c) == (int(g["num_x"]), int(g["num_c"]))
    np.testing.assert_array_equal(it.scaling.V, g["V"])
    np.testing.assert_array_equal(it.scaling.r, g["r"])
    assert_guess_equal(name, it.guess_x_tilde, g["guess_x"])
    assert_guess_equal(name, it.x_bnd_l, g["x_bnd_l"])
    assert_guess_equal(name, it.x_bnd_u, g["x_bnd_u"])
    it.scaling.w, it.scaling.W_ocp = float(g["w"]), g["W_ocp"].copy()
    np.testing.assert_array_equal(it.scaling.W, g["W"])
    np.testing.assert_array_equal(it.c_bnd_l, g["c_bnd_l"])
    np.testing.assert_array_equal(it.c_bnd_u, g["c_bnd_u"])


# --------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_backend_matches_reference(name, cuda_device):
    """The drop-in itself: ``OptimalControlProblem.initialise()`` with
    ``backend="cuda"`` (guess interpolation on the device, engine compiled, J / c
    scaling derived from g and the row norms of G on the device) and then the
    reference's ``evaluate_*`` surface, against the reference's own results."""
    g = golden(name)
    ocp = build_golden_problem(name)
    ocp.initialise()
    be = ocp._backend
    it = be.mesh_iterations[0]
    assert_guess_equal(name, it.guess_x_tilde, g["guess_x"])          # pcx_interp_guess
    assert_guess_equal(name, it.x_bnd_l, g["x_bnd_l"])                # pcx_expand_bounds
    assert_guess_equal(name, it.x_bnd_u, g["x_bnd_u"])
    # N1: w and W from the device (scaling.py:346-430).  Delta III's guess is a
    # singular point of its dynamics (the reference's own G is not finite there)
    if not bool(g["singular_guess"]):
        assert strict_err([it.scaling.w], [float(g["w"])]) <= RTOL
        e_W = strict_err(it.scaling.W_ocp, g["W_ocp"])
        print(f"cuda {name}: W_ocp {e_W:.1e}")
        assert e_W <= RTOL
    rows, cols = be.evaluate_G_structure()
    np.testing.assert_array_equal(rows, g["G_row"])
    np.testing.assert_array_equal(cols, g["G_col"])
    assert be.evaluate_G_num_nonzero() == int(g["G_nnz"])
    rows, cols = be.evaluate_H_structure()
    np.testing.assert_array_equal(rows, g["H_row"])
    np.testing.assert_array_equal(cols, g["H_col"])
    # exactly the reference's scaling for the value comparison
    it.scaling.w, it.scaling.W_ocp = float(g["w"]), g["W_ocp"].copy()
    it.push_scaling()
    np.testing.assert_array_equal(it.c_bnd_l, g["c_bnd_l"])
    np.testing.assert_array_equal(it.c_bnd_u, g["c_bnd_u"])
    _check_values(name, g, lambda x, lam, sg: dict(
        J=be.evaluate_J(x), g=be.evaluate_g(x), c=be.evaluate_c(x), dy=be.evaluate_dy(x),
        G_data=be.evaluate_G_nonzeros(x), H_data=be.evaluate_H_nonzeros(x, sg, lam)), "cuda")


@pytest.mark.gpu
@pytest.mark.parametrize("name", SMALL[:4])
def test_c_abi_structure_matches_reference(name, cuda_device):
    """``pcx_structure_jac`` / ``pcx_structure_hess`` (include/pcx.h) hand a C/C++
    host the same patterns, plus the row-major / lower-triangular views of the
    cyipopt contract (``pycollo/nlp.py:36-76``)."""
    g = golden(name)
    ocp = build_golden_problem(name)
    ocp.initialise()
    eng = ocp._backend.mesh_iterations[0].create_engine()
    rows, cols, perm = eng.structure_jac("ccs")
    np.testing.assert_array_equal(rows, g["G_row"])
    np.testing.assert_array_equal(cols, g["G_col"])
    np.testing.assert_array_equal(perm, np.arange(len(rows)))
    rows, cols, perm = eng.structure_jac("row_major")
    order = np.lexsort((g["G_col"], g["G_row"]))
    np.testing.assert_array_equal(rows, g["G_row"][order])
    np.testing.assert_array_equal(cols, g["G_col"][order])
    np.testing.assert_array_equal(perm, order)
    rows, cols, perm = eng.structure_hess("triu_ccs")
    np.testing.assert_array_equal(rows, g["H_row"])
    np.testing.assert_array_equal(cols, g["H_col"])
    rows, cols, perm = eng.structure_hess("tril_row_major")
    np.testing.assert_array_equal(rows, g["H_col"])
    np.testing.assert_array_equal(cols, g["H_row"])
    np.testing.assert_array_equal(perm, np.arange(len(rows)))

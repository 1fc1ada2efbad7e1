"""Host mesh/quadrature tables vs the reference's own (golden fixtures).

Fixtures in tests/golden/ were produced by executing pycollo/quadrature.py and
pycollo/mesh.py (oracle/make_golden.py), orders 2..20.  The default tables
(``Quadrature(method)``, ``tables="reference"``) reproduce the reference's
numerics -- numpy's Legendre roots, the closed-form weights and the
``numpy.linalg.solve`` of its simplifying-condition system -- and must equal the
fixtures TO THE LAST BIT, including the digits that ill-conditioned solve loses at
high order.  ``tables="exact"`` (collocation integrals by Gauss quadrature) is the
mathematically exact alternative; its distance from the reference grows with the
order and is bounded here.
"""
import numpy as np
import pytest

from helpers import GOLDEN, golden_mesh
from pycollo_b200.mesh import PhaseMesh, PhaseMeshData
from pycollo_b200.quadrature import Quadrature

TOL_A = {2: 1e-15, 3: 1e-15, 4: 2e-15, 5: 4e-15, 6: 6e-15, 7: 2e-14, 8: 2e-13, 9: 4e-13,
         10: 3e-12}


@pytest.mark.parametrize("method", ["lobatto", "radau"])
def test_default_tables_are_the_references_bit_for_bit(method):
    g = np.load(f"{GOLDEN}/quadrature_{method}.npz")
    q = Quadrature(method)
    for n in range(2, 21):
        np.testing.assert_array_equal(q.quadrature_point(n), g[f"points_{n}"])
        np.testing.assert_array_equal(q.quadrature_weight(n), g[f"weights_{n}"])
        np.testing.assert_array_equal(q.butcher_array(n), g[f"butcher_{n}"])
        np.testing.assert_array_equal(q.A_matrix(n), g[f"A_{n}"])
        np.testing.assert_array_equal(q.D_matrix(n), g[f"D_{n}"])
        assert q.A_matrix(n).shape == (n - 1, n)


@pytest.mark.parametrize("method", ["lobatto", "radau"])
def test_exact_tables_close_to_reference(method):
    g = np.load(f"{GOLDEN}/quadrature_{method}.npz")
    q = Quadrature(method, tables="exact")
    for n in range(2, 11):
        np.testing.assert_allclose(q.quadrature_point(n), g[f"points_{n}"], atol=4e-15, rtol=0)
        np.testing.assert_allclose(q.quadrature_weight(n), g[f"weights_{n}"], atol=1e-13, rtol=0)
        np.testing.assert_allclose(q.A_matrix(n), g[f"A_{n}"], atol=TOL_A[n], rtol=0)


def test_lobatto_weights_exact_low_order():
    """tests/unit/test_quadrature.py:48-55 of the reference."""
    q = Quadrature("lobatto")
    np.testing.assert_array_equal(q.quadrature_weight(2), np.array([0.5, 0.5]))
    np.testing.assert_allclose(q.quadrature_weight(3),
                               np.array([1 / 6, 2 / 3, 1 / 6]), rtol=0, atol=1e-16)


def test_radau_quirks_are_kept():
    """Radau: padding weight and last integration column exactly zero; weights sum
    to 2 (Lobatto: 1) -- pycollo/quadrature.py:116-139, SURVEY.md §8 a10."""
    q = Quadrature("radau")
    for n in range(2, 11):
        assert q.quadrature_weight(n)[-1] == 0.0
        assert np.all(q.A_matrix(n)[:, -1] == 0.0)
        assert abs(q.quadrature_weight(n).sum() - 2.0) < 1e-13
    ql = Quadrature("lobatto")
    for n in range(2, 11):
        assert abs(ql.quadrature_weight(n).sum() - 1.0) < 1e-13


def test_gauss_unsupported():
    with pytest.raises(ValueError):
        Quadrature("gauss")


@pytest.mark.parametrize("method", ["lobatto", "radau"])
@pytest.mark.parametrize("tag", ["uniform10x4", "ragged6"])
def test_mesh_matches_reference(method, tag):
    g = golden_mesh(f"mesh_{method}_{tag}")
    q = Quadrature(method)
    m = PhaseMeshData(q, PhaseMesh(len(g["N_K"]), g["sizes"], g["N_K"]), 2, 10)
    assert m.N == g["N"]
    np.testing.assert_array_equal(m.tau, g["tau"])             # bit for bit
    np.testing.assert_array_equal(m.h_K, g["h_K"])
    np.testing.assert_array_equal(m.W_matrix, g["W"])
    sI, sA = m.sI_matrix, m.sA_matrix
    assert np.array_equal(sI.indptr, g["sI"].indptr) and np.array_equal(sI.indices, g["sI"].indices)
    np.testing.assert_array_equal(sI.data, g["sI"].data)
    assert (sA != g["sA"]).nnz == 0


def test_mesh_adopts_reference_arrays_bit_exactly():
    """Drop-in path: the reference's own sI_matrix / W_matrix are consumed as given."""
    g = golden_mesh("mesh_radau_ragged6")
    q = Quadrature("radau")
    m = PhaseMeshData.from_reference_csr(q, g["N_K"], g["tau"], g["sI"], g["W"])
    assert np.array_equal(m.sI_matrix.toarray(), g["sI"].toarray())
    assert np.array_equal(m.W_matrix, g["W"])


def test_mesh_validation_errors():
    q = Quadrature("lobatto")
    with pytest.raises(ValueError):
        PhaseMesh(3, [0.5, 0.5], 4)
    with pytest.raises(ValueError):
        PhaseMesh(3, None, [4, 4])
    with pytest.raises(ValueError):
        PhaseMeshData(q, PhaseMesh(3, None, 3), 4, 10)
    with pytest.raises(ValueError):
        PhaseMeshData(q, PhaseMesh(3, None, 12), 4, 10)

"""Pin the oracle against every known answer the reference's tests hold for the
callback path: tests/unit/test_iteration.py:290-385 and
tests/unit/test_iteration_scaling.py of the reference.  The meshes used here are
the reference's OWN matrices (golden fixtures produced by executing
pycollo/mesh.py), so the pins check the restated algebra, not the mesh code.
"""
import numpy as np
import pytest

from helpers import GOLDEN, golden_mesh
from examples import problems as examples
from pycollo_b200.symbolic import build_ir


@pytest.fixture(scope="module")
def br():
    from oracle.blockwise import BlockwiseNLP
    from oracle.expand import ExpandedNLP
    ocp = examples.brachistochrone()
    ir = build_ir(ocp)
    mesh = golden_mesh("mesh_lobatto_uniform10x4")
    E = ExpandedNLP(ocp, ir.full_bounds, [mesh])
    B = BlockwiseNLP(ocp, ir.full_bounds, [mesh])
    g = np.load(f"{GOLDEN}/iteration_scaling_brachistochrone.npz")
    return E, B, g


def test_brachistochrone_sizes(br):
    E, B, g = br
    assert E.num_x == B.num_x == 125          # test_iteration.py:189
    assert E.num_c == B.num_c == 90           # test_iteration.py:383


def test_brachistochrone_scaling_vectors(br):
    """EXPECT_V_BR / EXPECT_R_BR (OCP-level values expanded on the mesh)."""
    E, B, g = br
    V = np.concatenate([np.repeat(B.V[:4], 31), B.V[4:]])
    r = np.concatenate([np.repeat(B.r[:4], 31), B.r[4:]])
    np.testing.assert_allclose(V, g["V"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(r, g["r"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(V * g["x_tilde"] + r, g["x"], atol=1e-6)


def test_brachistochrone_J(br):
    E, B, g = br
    for nlp in (E, B):
        np.testing.assert_almost_equal(nlp.J(g["x_tilde"]), 0.8243386694458454)  # :317-318


def test_brachistochrone_g(br):
    E, B, g = br
    expect = np.zeros(125)
    expect[124] = 10                                                  # :351-354
    for nlp in (E, B):
        np.testing.assert_allclose(nlp.g(g["x_tilde"]), expect)


def test_brachistochrone_c(br):
    E, B, g = br
    for nlp in (E, B):
        np.testing.assert_allclose(nlp.c(g["x_tilde"]), np.zeros(90), atol=10e-2)  # :383-385
        assert np.max(np.abs(nlp.c(g["x_tilde"]))) < 1e-7     # converged solution


def test_double_pendulum_pins():
    from oracle.blockwise import BlockwiseNLP
    ocp = examples.double_pendulum()
    ir = build_ir(ocp)
    B = BlockwiseNLP(ocp, ir.full_bounds, [golden_mesh("mesh_lobatto_uniform10x4")])
    g = np.load(f"{GOLDEN}/iteration_scaling_double_pendulum.npz")
    assert B.num_x == 190 and B.num_c == 121                        # :207, 222
    assert B.J(g["x_tilde"]) == 100                                 # :302
    expect = np.zeros(190)
    expect[186] = 1000                                              # :333-336
    np.testing.assert_allclose(B.g(g["x_tilde"]), expect)
    offs = np.concatenate([np.repeat(B.V[:6], 31), B.V[6:]])
    np.testing.assert_allclose(offs, g["V"], rtol=1e-15)
    np.testing.assert_allclose(1.0 / offs, g["V_inv"], rtol=1e-14)


def test_two_restatements_agree_and_match_finite_differences(br):
    """G and H have no reference pins: the literal symbolic
    expansion and the blockwise restatement must agree on pattern and values,
    and both must match central differences of c and of G^T lam."""
    E, B, g = br
    assert np.array_equal(E.G_structure()[0], B.G_structure()[0])
    assert np.array_equal(E.G_structure()[1], B.G_structure()[1])
    assert np.array_equal(E.H_structure()[0], B.H_structure()[0])
    assert np.array_equal(E.H_structure()[1], B.H_structure()[1])
    rows, cols = B.G_structure()
    assert np.all(np.diff(cols * B.num_c + rows) > 0)             # CCS order, no duplicates
    hr, hc = B.H_structure()
    assert np.all(hr <= hc) and np.all(np.diff(hc * B.num_x + hr) > 0)
    rng = np.random.default_rng(3)
    x = rng.uniform(-0.5, 0.5, B.num_x)
    lam = rng.standard_normal(B.num_c)
    np.testing.assert_allclose(E.G_nonzeros(x), B.G_nonzeros(x), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(E.H_nonzeros(x, 0.7, lam), B.H_nonzeros(x, 0.7, lam),
                               rtol=1e-10, atol=1e-11)
    G = np.zeros((B.num_c, B.num_x))
    G[rows, cols] = B.G_nonzeros(x)
    H = np.zeros((B.num_x, B.num_x))
    H[hr, hc] = B.H_nonzeros(x, 0.7, lam)
    H = H + np.triu(H, 1).T
    eps = 1e-6
    for j in rng.choice(B.num_x, 12, replace=False):
        d = np.zeros(B.num_x)
        d[j] = eps
        fd = (B.c(x + d) - B.c(x - d)) / (2 * eps)
        np.testing.assert_allclose(G[:, j], fd, atol=2e-6, rtol=1e-6)

        def grad_L(z):
            Gz = np.zeros((B.num_c, B.num_x))
            Gz[rows, cols] = B.G_nonzeros(z)
            return 0.7 * B.g(z) + Gz.T @ lam
        fdh = (grad_L(x + d) - grad_L(x - d)) / (2 * eps)
        np.testing.assert_allclose(H[:, j], fdh, atol=5e-4, rtol=1e-5)


@pytest.mark.parametrize("method", ["lobatto", "radau"])
def test_restatements_agree_on_cart_pole_small(method):
    """brute-force expansion of cart-pole on K=2 x 4 nodes gives
    nnz_G / nnz_H(tri) = 236/35 (Lobatto) and 189/30 (Radau, zeros pruned)."""
    from helpers import make_meshes, oracle_meshes
    from oracle.blockwise import BlockwiseNLP
    from oracle.expand import ExpandedNLP
    ocp = examples.cart_pole_swing_up(quadrature_method=method)
    ir = build_ir(ocp)
    meshes = oracle_meshes(make_meshes(ocp, method, 2, 4))
    E = ExpandedNLP(ocp, ir.full_bounds, meshes)
    B = BlockwiseNLP(ocp, ir.full_bounds, meshes)
    expect = (236, 35) if method == "lobatto" else (189, 30)
    assert (len(E.G_rows), len(E.H_rows)) == expect
    assert (len(B.G_structure()[0]), len(B.H_structure()[0])) == expect
    rng = np.random.default_rng(0)
    x = rng.uniform(-0.5, 0.5, B.num_x)
    lam = rng.standard_normal(B.num_c)
    np.testing.assert_allclose(E.c(x), B.c(x), rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(E.G_nonzeros(x), B.G_nonzeros(x), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(E.H_nonzeros(x, 1.0, lam), B.H_nonzeros(x, 1.0, lam),
                               rtol=1e-10, atol=1e-11)
    if method == "radau":
        Bk = BlockwiseNLP(ocp, ir.full_bounds, meshes, prune=False)
        Ek = ExpandedNLP(ocp, ir.full_bounds, meshes, prune=False)
        assert (len(Bk.G_structure()[0]), len(Bk.H_structure()[0])) == (236, 35)
        assert (len(Ek.G_rows), len(Ek.H_rows)) == (236, 35)


def test_lowered_equations_match_the_reference_known_answers():
    """The reference's own known-answer tests for the substituted problem functions
    (``tests/unit/test_backend_casadi.py:1328-1402`` double pendulum -- phase-level
    auxiliary data shadow problem-level ones: g = -9.81, k1 = 1/12 -- ``:1405-1435``
    brachistochrone, ``:1470-1490`` integrand, ``:1521-1566`` objectives) restated in the
    user basis and compared at random points with what the lowering hands the code
    generator (``symbolic.build_ir``)."""
    import sympy as sym
    from examples import problems
    from pycollo_b200.symbolic import build_ir
    rng = np.random.default_rng(12)

    def same(ours, expect, syms, n=6):
        f = sym.lambdify(syms, [sym.sympify(e) for e in ours], "numpy")
        g = sym.lambdify(syms, [sym.sympify(e) for e in expect], "numpy")
        for _ in range(n):
            pt = rng.uniform(0.3, 1.4, len(syms))
            np.testing.assert_allclose(np.array(f(*pt), dtype=float), np.array(g(*pt), dtype=float),
                                       rtol=1e-13, atol=1e-13)

    ir = build_ir(problems.double_pendulum())
    ph = ir.phases[0]
    a0, a1, v0, v1 = ph.y
    T0, T1 = ph.u
    m0, p0 = ir.s
    g_, d0, k0, m1, p1, k1 = -9.81, 0.5, sym.Rational(1, 12), 1.0, 0.5, sym.Rational(1, 12)
    l0 = p0 + d0
    I0, I1 = m0 * (k0 ** 2 + p0 ** 2), m1 * (k1 ** 2 + p1 ** 2)
    c0, s0, c1, s1 = sym.cos(a0), sym.sin(a0), sym.cos(a1), sym.sin(a1)
    M00, M01, M11 = I0 + m1 * l0 ** 2, m1 * p1 * l0 * (s0 * s1 + c0 * c1), I1
    M10 = M01
    K0 = T0 + g_ * (m0 * p0 + m1 * l0) * c0 + m1 * p1 * l0 * (s1 * c0 - s0 * c1) * v1 ** 2
    K1 = T1 + g_ * m1 * p1 * c1 + m1 * p1 * l0 * (s0 * c1 - s1 * c0) * v0 ** 2
    detM = M00 * M11 - M01 * M10
    expect = (v0, v1, (M11 * K0 - M01 * K1) / detM, (M00 * K1 - M10 * K0) / detM)
    assert len(ph.f) == 4 and len(ph.p) == 0 and len(ph.g) == 1
    same(ph.f, expect, [a0, a1, v0, v1, T0, T1, m0, p0])
    same(ph.g, (T0 ** 2 + T1 ** 2,), [T0, T1])
    assert str(ir.J) == "q0_P0"                                   # J = the integral variable

    ir = build_ir(problems.brachistochrone())
    ph = ir.phases[0]
    (x, y, v), (u,) = ph.y, ph.u
    same(ph.f, (v * sym.sin(u), v * sym.cos(u), 9.81 * sym.cos(u)), [x, y, v, u])
    assert str(ir.J) == "tF_P0"                                   # J = the final time

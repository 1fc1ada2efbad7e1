"""Sparsity structure (bit-exact) and device tables (through the CPU emulation)
against the oracle, on the edge cases the domain has: ragged meshes, sections of
2..20 nodes, Radau zero pruning, free/fixed times, static parameters, tiles
smaller than the mesh, random objective/constraint scaling."""
import numpy as np
import pytest

from emulator import emulate
from helpers import RAGGED_NODES, RAGGED_SIZES, build_case, max_err
from pycollo_b200 import engine as E
from examples import problems as examples

CASES = [
    ("brachistochrone", 10, 4, None, {}),
    ("cart_pole_swing_up", 12, 4, None, dict(max_tile_nodes=16)),
    ("hypersensitive", 6, RAGGED_NODES, RAGGED_SIZES, dict(max_tile_nodes=20)),
    ("double_pendulum", 6, RAGGED_NODES, RAGGED_SIZES, dict(max_tile_nodes=20)),
    ("free_flying_robot", 5, [4, 6, 3, 5, 4], None, dict(max_tile_nodes=12)),
    ("multiphase_sliding_mass", 4, 4, None, {}),
    ("space_shuttle_reentry", 4, [4, 3, 5, 4], None, dict(max_tile_nodes=10)),
    # sections of 17-20 nodes (Settings.collocation_points_max = 20): the recipe word's
    # node field is 5 bits wide since round 2
    ("brachistochrone", 4, [18, 17, 20, 19], None, {}),
    ("hypersensitive", 3, [20, 4, 18], [0.5, 0.2, 0.3], dict(max_tile_nodes=24)),
]


def test_delta_iii_four_phases_patterns_and_values():
    """BASELINE config 4 problem: 4 phases, 18 linkage constraints, phase-dependent
    auxiliary data, path constraints."""
    low, B, scal = build_case(examples.delta_iii_launch_vehicle(), "lobatto", 3, 4, seed=2)
    S = low.S
    assert S.P == 4 and S.NB == 18
    assert np.array_equal(S.G_structure()[0], B.G_structure()[0])
    assert np.array_equal(S.G_structure()[1], B.G_structure()[1])
    assert np.array_equal(S.H_structure()[0], B.H_structure()[0])
    assert np.array_equal(S.H_structure()[1], B.H_structure()[1])
    tabs = E.build_tables(S, low.layouts)
    st = E.scaling_tables(S, low.layouts, *scal)
    rng = np.random.default_rng(1)
    x = rng.uniform(-0.5, 0.5, S.num_x)
    lam = rng.standard_normal(S.num_c)
    out = emulate(low, tabs, st, x, lam, 0.7, flags=63)
    for key, ref in (("jac", B.G_nonzeros(x)), ("hess", B.H_nonzeros(x, 0.7, lam)),
                     ("c", B.c(x)), ("grad", B.g(x)), ("dy", B.dy(x))):
        assert max_err(out[key], ref) <= 1e-12, key


@pytest.mark.parametrize("method", ["lobatto", "radau"])
@pytest.mark.parametrize("name,K,nodes,sizes,kw", CASES)
def test_patterns_are_bit_identical_to_oracle(name, method, K, nodes, sizes, kw):
    low, B, _ = build_case(getattr(examples, name)(), method, K, nodes, sizes, **kw)
    S = low.S
    gr, gc = S.G_structure()
    br, bc = B.G_structure()
    assert S.nnz_g == len(br)
    assert np.array_equal(gr, br) and np.array_equal(gc, bc)
    hr, hc = S.H_structure()
    bhr, bhc = B.H_structure()
    assert S.nnz_h == len(bhr)
    assert np.array_equal(hr, bhr) and np.array_equal(hc, bhc)
    assert (S.num_x, S.num_c) == (B.num_x, B.num_c)


@pytest.mark.parametrize("method", ["lobatto", "radau"])
@pytest.mark.parametrize("name,K,nodes,sizes,kw", CASES)
def test_device_tables_reproduce_oracle_values(name, method, K, nodes, sizes, kw):
    low, B, scal = build_case(getattr(examples, name)(), method, K, nodes, sizes,
                              seed=4, **kw)
    S = low.S
    tabs = E.build_tables(S, low.layouts)
    st = E.scaling_tables(S, low.layouts, *scal)
    rng = np.random.default_rng(5)
    x = rng.uniform(-0.5, 0.5, S.num_x)
    lam = rng.standard_normal(S.num_c)
    out = emulate(low, tabs, st, x, lam, 0.7, flags=63)
    tol = 1e-12
    assert max_err([out["f"]], [B.J(x)]) <= tol
    assert max_err(out["grad"], B.g(x)) <= tol
    assert max_err(out["c"], B.c(x)) <= tol
    assert max_err(out["dy"], B.dy(x)) <= tol
    assert max_err(out["jac"], B.G_nonzeros(x)) <= tol
    assert max_err(out["hess"], B.H_nonzeros(x, 0.7, lam)) <= tol


def test_unpruned_radau_pattern_switch():
    ocp = examples.cart_pole_swing_up()
    ocp.settings.prune_zero_quadrature_coefficients = False
    low, B, _ = build_case(ocp, "radau", 10, 4)
    assert low.S.prune is False
    assert (low.S.nnz_g, low.S.nnz_h) == (len(B.G_structure()[0]), len(B.H_structure()[0]))
    ocp2 = examples.cart_pole_swing_up()
    low2, _, _ = build_case(ocp2, "lobatto", 10, 4)
    assert (low.S.nnz_g, low.S.nnz_h) == (low2.S.nnz_g, low2.S.nnz_h)   # Lobatto == unpruned


def test_config2_counts_and_tiling():
    """BASELINE config 2: cart-pole at 10^5 nodes."""
    low, _, _ = build_case(examples.cart_pole_swing_up(), "lobatto", 33333, 4, oracle=False)
    S = low.S
    N = 100000
    assert (S.num_x, S.num_c) == (500001, 399997)
    assert S.nnz_g == 38 * (N - 1) + N + 1 == 3899963
    assert S.nnz_h == 5 * N == 500000
    # one wave of 148 x 6 CTAs: 887 tiles + the border CTA
    assert (S.num_tiles + 1) % 148 == 0 and S.max_tile_nodes <= S.threads
    # tiles partition the sections; runs partition the tiles
    assert S.tile_k0[0] == 0 and S.tile_k1[-1] == 33333
    assert np.array_equal(S.tile_k0[1:], S.tile_k1[:-1])
    d = S.tile_desc
    assert np.array_equal(d[1:, 5], d[:-1, 6]) and d[-1, 6] == len(S.run_slo)
    assert S.max_tile_runs <= 2


def test_recipe_word_limits_are_enforced():
    import sympy as sym
    from pycollo_b200 import OptimalControlProblem
    ys = sym.symbols("y0:70")
    ocp = OptimalControlProblem("wide")
    ph = ocp.new_phase("A")
    ph.state_variables = ys
    ph.state_equations = [-y for y in ys]
    ph.bounds.initial_time, ph.bounds.final_time = 0, 1
    ph.bounds.state_variables = [[-1, 1]] * 70
    ph.guess.time = [0, 1]
    ph.guess.state_variables = [[0, 0]] * 70
    ocp.objective_function = ph.final_state_variables[0]
    with pytest.raises(ValueError, match="recipe word"):
        build_case(ocp, "lobatto", 4, 4, oracle=False)

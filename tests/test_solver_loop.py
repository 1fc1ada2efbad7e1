"""Row R1 stand-in: a host NLP solver consumes the callbacks end to end.

The reference's solve (``pycollo/backend.py:1807-1827``) is ``ca.nlpsol`` -> IPOPT;
neither exists in this image.  ``tests/solver_loop.py`` is an interior-point Newton
method over the cyipopt callback contract (``pycollo/nlp.py:36-76``).  Solving with
it checks that G and H are jointly right -- the iterates converge to the
objectives the reference's integration tests pin
(``tests/integration/test_brachistochrone.py:159-166`` 0.82434 rtol 1e-4,
``tests/unit/test_iteration.py:317-318`` 0.8243386694458454 on the initial mesh,
``tests/integration/test_multiphase.py:78-84`` 0.4472136 rtol 1e-4) -- and that the
CUDA callbacks and the CPU oracle drive the solver through the same iterates
(identical iteration counts).  IPOPT's own iteration counts remain "not run".
"""
import numpy as np
import pytest

from examples import problems
from solver_loop import OracleCallbacks, solve

CASES = {
    # name: (builder, kwargs, objective pin, rtol)
    "brachistochrone": ("brachistochrone", {}, 0.8243386694458454, 1e-8),
    "multiphase": ("multiphase_sliding_mass", dict(num_phases=3), 0.4472136, 1e-4),
    "cart_pole": ("cart_pole_swing_up", {}, None, None),
}


def _oracle_for(ocp, it):
    from oracle.blockwise import BlockwiseNLP
    return BlockwiseNLP(
        ocp, it.low.ir.full_bounds,
        [dict(N=m.N, sI=m.sI_matrix, sA=m.sA_matrix, W=m.W_matrix) for m in it.mesh.p],
        W_ocp=it.scaling.W_ocp, w=it.scaling.w, prune=it.S.prune,
        scaling_method=ocp.settings.scaling_method)


def _solve_with(cb, it):
    res = solve(cb, it.guess_x_tilde, it.x_bnd_l, it.x_bnd_u, it.c_bnd_l, it.c_bnd_u)
    res.J_user = it.scaling.unscale_J(res.fun)
    return res


@pytest.mark.parametrize("name", ["brachistochrone", "multiphase"])
def test_oracle_callbacks_reach_the_reference_objective(name):
    builder, kw, pin, rtol = CASES[name]
    ocp = getattr(problems, builder)(**kw)
    ocp.settings.defer_engine = True
    ocp.initialise()
    it = ocp._backend.mesh_iterations[0]
    res = _solve_with(OracleCallbacks(_oracle_for(ocp, it)), it)
    print(f"oracle {name}: J = {res.J_user:.13f} in {res.nit} iterations, "
          f"constraint violation {res.constr_violation:.1e}")
    assert res.success and res.constr_violation <= 1e-8
    np.testing.assert_allclose(res.J_user, pin, rtol=rtol)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_callbacks_solve_like_the_oracle(name, cuda_device):
    builder, kw, pin, rtol = CASES[name]
    ocp = getattr(problems, builder)(**kw)
    ocp.initialise()                                   # engine, scaling from the device
    it = ocp._backend.mesh_iterations[0]
    cb = ocp._backend.nlp_callbacks("cyipopt")
    res = _solve_with(cb, it)
    ref = _solve_with(OracleCallbacks(_oracle_for(ocp, it)), it)
    print(f"cuda {name}: J = {res.J_user:.13f} in {res.nit} iterations (oracle callbacks: "
          f"{ref.nit}), |dJ| = {abs(res.J_user - ref.J_user):.1e}, x uploads "
          f"{cb.num_x_uploads}, launches {cb.num_launches}")
    assert res.success and res.constr_violation <= 1e-8
    assert res.nit == ref.nit                            # same iterates
    np.testing.assert_allclose(res.J_user, ref.J_user, rtol=1e-9)
    np.testing.assert_allclose(res.x, ref.x, atol=1e-7)
    if pin is not None:
        np.testing.assert_allclose(res.J_user, pin, rtol=rtol)


def test_backend_solve_nlp_runs_the_builtin_solver():
    """``Cuda.solve_nlp`` (``backend.py:1807-1827``): guess, bounds, tolerance and iteration
    limit of the current mesh iteration handed to the built-in interior-point solver;
    an ``NlpResult`` keyed like nlpsol's output comes back.  CPU: the callbacks are the
    oracle's (the engine is deferred); the GPU test above drives the same solver with
    the CUDA callbacks."""
    ocp = problems.brachistochrone()
    ocp.settings.defer_engine = True
    ocp.initialise()
    backend = ocp._backend
    it = backend.mesh_iterations[0]
    res = backend.solve_nlp(callbacks=OracleCallbacks(_oracle_for(ocp, it)))
    assert res.info["solver"] == "pycollo_b200.ipnewton" and res.info["success"]
    assert res.info["kkt_error"] <= ocp.settings.nlp_tolerance
    assert res.solve_time > 0
    sol = res.solution
    assert sol["x"].shape == (it.S.num_x,) and sol["lam_g"].shape == (it.S.num_c,)
    assert sol["g"].shape == (it.S.num_c,) and np.max(np.abs(sol["g"])) <= 1e-8
    np.testing.assert_allclose(it.scaling.unscale_J(sol["f"]), 0.8243386694458454, rtol=1e-8)
    assert np.all(sol["x"] >= it.x_bnd_l - 1e-12) and np.all(sol["x"] <= it.x_bnd_u + 1e-12)

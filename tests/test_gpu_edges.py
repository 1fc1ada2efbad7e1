"""GPU edge cases: smallest meshes, one-node-per-thread limits, batched mesh
error, C-ABI error behaviour of the sharding / mesh-error entry points."""
import numpy as np
import pytest

from helpers import build_case, make_engine, max_err
from pycollo_b200 import engine as E
from examples import problems as examples

pytestmark = pytest.mark.gpu
ALL = E.EVAL_C | E.EVAL_DY | E.EVAL_JAC | E.EVAL_HESS | E.EVAL_F | E.EVAL_GRAD


def _parity(name, method, K, nodes, sizes=None, **kw):
    low, B, scal = build_case(getattr(examples, name)(), method, K, nodes, sizes, seed=5, **kw)
    eng = make_engine(low, scal)
    rng = np.random.default_rng(3)
    x = rng.uniform(-0.5, 0.5, low.S.num_x)
    lam = rng.standard_normal(low.S.num_c)
    out = eng.eval_host(ALL, x, lam, 1.3)
    ref = dict(f=[B.J(x)], grad=B.g(x), c=B.c(x), dy=B.dy(x), jac=B.G_nonzeros(x),
               hess=B.H_nonzeros(x, 1.3, lam))
    for k, r in ref.items():
        got = out[k] if k == "f" else out[k][0]
        assert max_err(got, r) <= 1e-12, (k, low.S.threads, low.S.num_tiles)
    return low


@pytest.mark.parametrize("method", ["lobatto", "radau"])
@pytest.mark.parametrize("name", ["brachistochrone", "double_pendulum", "multiphase_sliding_mass"])
def test_smallest_meshes(name, method):
    """N = 3 is the smallest supported mesh (one interior node): as one 3-node
    section and as two 2-node sections; N = 2 is refused with a clear error."""
    low = _parity(name, method, 1, 3)
    assert all(t.N == 3 for t in low.S.ph)
    low = _parity(name, method, 2, 2)
    assert all(t.N == 3 and t.K == 2 for t in low.S.ph)
    with pytest.raises(ValueError, match="at least 3 mesh nodes"):
        build_case(getattr(examples, name)(), method, 1, 2, oracle=False)


@pytest.mark.parametrize("nodes", [3, 10, 16, 17, 20])
def test_single_section_orders(nodes):
    """Up to Settings.collocation_points_max = 20 nodes per section (the recipe word's
    node field was 4 bits wide in round 1: sections of 17-20 nodes scattered wrong
    Jacobian values silently)."""
    _parity("cart_pole_swing_up", "lobatto", 1, nodes)
    _parity("hypersensitive", "radau", 1, nodes)


def test_sections_of_17_to_20_nodes_in_one_mesh():
    _parity("brachistochrone", "lobatto", 4, [18, 17, 20, 19])
    _parity("double_pendulum", "radau", 3, [20, 4, 18], max_tile_nodes=24)
    # the ph mesh of a 20-node section has 21 nodes: the mesh-error engine's limit
    with pytest.raises(ValueError):
        _parity("hypersensitive", "lobatto", 1, 33)


def test_cta_sizes_follow_the_mesh():
    """32 / 64 threads per CTA for meshes of <= 32 / <= 64 nodes; larger meshes: 192
    for small expression bodies (cart-pole), 128 for large ones (robot); a tile never
    holds more nodes than threads."""
    for name, K, want in (("cart_pole_swing_up", 10, 32), ("cart_pole_swing_up", 20, 64),
                          ("cart_pole_swing_up", 30, 192), ("free_flying_robot", 30, 128)):
        low = _parity(name, "lobatto", K, 4)
        assert low.S.threads == want and low.S.max_tile_nodes <= want
    # one tile per instance takes the single-CTA path (tile + border pass in one CTA)
    low = _parity("free_flying_robot", "lobatto", 10, 4)
    assert low.S.num_tiles == 1


def test_many_small_tiles_with_sections_as_large_as_a_tile():
    # sections of 10 nodes with tiles capped at 10 nodes: one section per tile
    _parity("space_shuttle_reentry", "lobatto", 12, 10, max_tile_nodes=10)
    _parity("double_pendulum", "radau", 9, [10, 2, 10, 2, 10, 2, 10, 2, 10], max_tile_nodes=11)


def test_batched_mesh_error_matches_single():
    from pycollo_b200.mesh import Mesh, PhaseMesh
    from pycollo_b200.mesh_refinement import MeshErrorEvaluator
    from pycollo_b200.quadrature import Quadrature
    ocp = examples.free_flying_robot()
    ocp.settings.scaling_method = "none"
    mesh = Mesh(Quadrature("lobatto"), [PhaseMesh(7, None, [4, 5, 3, 6, 4, 2, 5])], 2, 16)
    one = MeshErrorEvaluator(ocp, mesh)
    many = MeshErrorEvaluator(ocp, mesh, batch=3)
    S = one.engine.S
    rng = np.random.default_rng(0)
    X = rng.uniform(-0.3, 0.3, (3, S.num_x))
    res = many(X)
    for b in range(3):
        single = one(X[b])
        for (a, r, m), (ab, rb, mb) in zip(single, res):
            assert np.array_equal(a, ab[b]) and np.array_equal(r, rb[b]) and np.array_equal(m, mb[b])
    assert res[0][2].shape == (3, 7)


def test_shard_and_mesh_error_argument_errors():
    low, _, scal = build_case(examples.cart_pole_swing_up(), "lobatto", 200, 4, oracle=False)
    eng = E.Engine(low.S, low.layouts, low.header)
    x = np.zeros(low.S.num_x)
    with pytest.raises(E.PcxError, match="pcx_set_scaling"):
        eng.mesh_error_host(x)                      # scaling not set yet
    eng.set_scaling(*scal)
    with pytest.raises(E.PcxError, match="tile range"):
        eng.set_shard(3, 2)
    with pytest.raises(E.PcxError, match="tile range"):
        eng.set_shard(0, low.S.num_tiles + 1)
    import torch
    xd = torch.zeros(low.S.num_x, dtype=torch.float64, device="cuda")
    jd = torch.zeros(low.S.nnz_g, dtype=torch.float64, device="cuda")
    with pytest.raises(E.PcxError, match="pcx_set_shard"):
        eng.apply_border(E.EVAL_JAC, xd, jac=jd)    # not sharded
    eng.set_shard(0, low.S.num_tiles)               # the full range is "not sharded" again
    out = eng.eval_host(E.EVAL_JAC, x)
    assert np.all(np.isfinite(out["jac"]))


@pytest.mark.parametrize("problem", ["double_pendulum", "multiphase_sliding_mass", "free_flying_robot"])
def test_guess_interpolation_matches_scipy_interp1d(problem):
    """Row N3: pcx_interp_guess vs the reference's own scipy call (oracle/guess.py),
    including abscissae a hair outside the previous mesh (extrapolation branch)."""
    from oracle.guess import interpolate_guess_to_mesh
    from pycollo_b200.mesh import Mesh, PhaseMesh
    from pycollo_b200.quadrature import Quadrature
    ocp = getattr(examples, problem)()
    low, _, scal = build_case(ocp, "lobatto", 9, [4, 6, 3, 8, 5, 4, 9, 2, 5],
                              [0.1, 0.15, 0.05, 0.2, 0.1, 0.12, 0.08, 0.1, 0.1], oracle=False)
    S = low.S
    eng = make_engine(low, scal)
    rng = np.random.default_rng(12)
    prev_tau, tau, ys, us, qs, ts, xprev = [], [], [], [], [], [], []
    for irp, t, m in zip(low.ir.phases, S.ph, low.meshes):
        M = int(rng.integers(5, 40))
        pt = np.sort(rng.uniform(-1, 1, M))
        pt[0], pt[-1] = -1.0, 1.0 - 1e-13          # last new node lies just outside
        prev_tau.append(pt)
        tau.append(np.asarray(m.tau))
        y, u = rng.standard_normal((irp.n_y, M)), rng.standard_normal((irp.n_u, M))
        q, tt = rng.standard_normal(irp.n_q), rng.standard_normal(irp.n_t)
        ys.append(y); us.append(u); qs.append(q); ts.append(tt)
        xprev += [y.ravel(), u.ravel(), q, tt]
    s = rng.standard_normal(S.NS)
    xprev.append(s)
    got = eng.interp_guess_host(np.concatenate(xprev), prev_tau, tau)
    ref = interpolate_guess_to_mesh(prev_tau, tau, ys, us, qs, ts, s)
    assert got.shape == ref.shape == (S.num_x,)
    assert np.max(np.abs(got - ref)) <= 1e-13 * (1.0 + np.abs(ref).max())


def test_cyipopt_adapter_orderings_and_staging():
    """NlpCallbacks (pycollo/nlp.py:36-76): row-major Jacobian / lower-triangular
    Hessian produced by the device-side permutation equal the engine's CCS values
    re-ordered on the host, for both orderings and repeated calls."""
    from pycollo_b200.backend import Cuda
    from pycollo_b200.nlp import NlpCallbacks
    ocp = examples.double_pendulum()
    examples.set_mesh(ocp, 12, 5)
    backend = Cuda(ocp)
    for step in ("create_bounds", "create_scaling", "create_quadrature",
                 "create_initial_mesh", "create_guess", "create_mesh_iterations"):
        getattr(backend, step)()
    it = backend.current_iteration
    it.generate_nlp()
    rng = np.random.default_rng(1)
    cb, cc = NlpCallbacks(it, "cyipopt"), NlpCallbacks(it, "casadi")
    for _ in range(3):
        x = it.guess_x_tilde + 0.1 * rng.standard_normal(it.num_x)
        lam = rng.standard_normal(it.num_c)
        g_ccs = backend.evaluate_G_nonzeros(x)
        h_ccs = backend.evaluate_H_nonzeros(x, 0.3, lam)
        assert np.array_equal(cb.jacobian(x), g_ccs[cb.g_perm])
        assert np.array_equal(cc.jacobian(x), g_ccs)
        # fused / single-output variants are separate compilations: rounding only
        assert max_err(cb.hessian(x, lam, 0.3), h_ccs[cb.h_perm]) <= 1e-13
        assert max_err(cc.hessian(x, lam, 0.3), h_ccs) <= 1e-13
    # iterate cache: the five callbacks of one solver iterate ship x once and launch
    # three kernels (f+grad+c fused, Jacobian, Hessian)
    x = it.guess_x_tilde + 0.05
    lam = rng.standard_normal(it.num_c)
    up0, l0 = cb.num_x_uploads, dict(cb.num_launches)
    f = cb.objective(x)
    g = cb.gradient(x).copy()
    c = cb.constraints(x).copy()
    jac = cb.jacobian(x).copy()
    hes = cb.hessian(x, lam, 1.0).copy()
    assert cb.num_x_uploads == up0 + 1
    assert [cb.num_launches[k] - l0[k] for k in ("point", "jacobian", "hessian")] == [1, 1, 1]
    cb.objective(x); cb.jacobian(x)                                # same iterate again: cached
    assert cb.num_x_uploads == up0 + 1 and cb.num_launches["jacobian"] == l0["jacobian"] + 1
    assert f == backend.evaluate_J(x)
    assert np.array_equal(g, backend.evaluate_g(x)) and np.array_equal(c, backend.evaluate_c(x))
    assert np.array_equal(jac, backend.evaluate_G_nonzeros(x)[cb.g_perm])
    assert max_err(hes, backend.evaluate_H_nonzeros(x, 1.0, lam)) <= 1e-13
    x2 = x.copy(); x2[-1] += 1e-9                                  # one entry differs: new iterate
    assert cb.objective(x2) == backend.evaluate_J(x2) and cb.num_x_uploads == up0 + 2
    cs = NlpCallbacks(it, "cyipopt", x_check="sampled")
    assert np.array_equal(cs.jacobian(x), jac) and cs.objective(x, new_x=False) == f
    rows, cols = cb.jacobianstructure()
    assert np.all(np.diff(rows) >= 0)                              # row-major
    hr, hc = cb.hessianstructure()
    assert np.all(hr >= hc)                                        # lower triangle
    # derivative_level = 1: no Hessian is generated, compiled or exposed
    ocp1 = examples.double_pendulum()
    ocp1.settings.derivative_level = 1
    ocp1.initialise()
    cb1 = ocp1._backend.nlp_callbacks()
    assert not hasattr(cb1, "hessian") and not hasattr(cb1, "hessianstructure")
    x1 = ocp1._backend.mesh_iterations[0].guess_x_tilde
    assert cb1.jacobian(x1).shape == (ocp1._backend.evaluate_G_num_nonzero(),)
    with pytest.raises(ValueError, match="derivative_level"):
        ocp1._backend.evaluate_H_nonzeros(x1, 1.0, None)


def test_backend_surface_callables():
    """The callables the rest of pycollo reaches for on a backend:
    ``g_iter_scale_callable([x; w])``, ``G_iter_scale_callable([x; W])``
    (``scaling.py:361-362, 392-394``), ``evaluate_dy_on_mesh`` (the hook replacing what
    ``mesh_refinement.py:109, 148-151`` reads) and ``process_solution``."""
    from pycollo_b200.backend import NlpResult
    ocp = examples.free_flying_robot()
    examples.set_mesh(ocp, 9, 4)
    ocp.initialise()
    b = ocp._backend
    it = b.current_iteration
    rng = np.random.default_rng(4)
    x = it.guess_x_tilde + 0.01 * rng.standard_normal(it.num_x)
    sc = it.scaling
    # the scales of the argument replace the engine's: linear in w / row-wise linear in W
    g = b.evaluate_g(x)
    assert max_err(b.g_iter_scale_callable(np.append(x, 3.0 * sc.w)), 3.0 * g) <= 1e-15
    G = b.evaluate_G(x).tocsr()
    W2 = sc.W_ocp * rng.uniform(0.5, 2.0, sc.W_ocp.size)
    G2 = b.G_iter_scale_callable(np.concatenate([x, W2])).tocsr()
    ratio = sc._expand_c_to_mesh(W2) / sc.W
    assert abs(G2 - G.multiply(ratio[:, None]).tocsr()).max() <= 1e-12 * abs(G).max()
    assert np.array_equal(b.G_iter_scale_callable(x).toarray(), G.toarray())
    # dy on a mesh given from outside, user basis: the iteration's own mesh reproduces dy
    dy = b.evaluate_dy_on_mesh(it.mesh, sc.unscale_x(x))
    assert max_err(dy, b.evaluate_dy(x)) <= 1e-12
    sol = b.process_solution(it, NlpResult(solution={"x": x, "f": b.evaluate_J(x)}, info=None,
                                           solve_time=0.0))
    assert sol.J == b.evaluate_J(x) and np.array_equal(sol.x, x)


def test_cyipopt_adapter_uploads_from_the_callers_arrays():
    """``register_inputs``: the caller's x / multiplier arrays are page-locked on first
    sight and become the DMA source (no staging copy).  Results equal the staging
    path's bit for bit; an array modified IN PLACE and flagged new is read again."""
    from pycollo_b200.nlp import NlpCallbacks
    ocp = examples.cart_pole_swing_up()
    examples.set_mesh(ocp, 3000, 4)
    ocp.initialise()
    it = ocp._backend.current_iteration
    rng = np.random.default_rng(2)
    x = it.guess_x_tilde + 0.05 * rng.standard_normal(it.num_x)
    lam = rng.standard_normal(it.num_c)
    assert x.nbytes >= (1 << 16)
    reg = NlpCallbacks(it, "cyipopt", register_inputs=True)
    stg = NlpCallbacks(it, "cyipopt", register_inputs=False)
    for rnd in range(3):
        f = reg.objective(x, new_x=True)
        assert f == stg.objective(x, new_x=True)
        assert np.array_equal(reg.gradient(x, new_x=False), stg.gradient(x, new_x=False))
        assert np.array_equal(reg.constraints(x, new_x=False), stg.constraints(x, new_x=False))
        assert np.array_equal(reg.jacobian(x, new_x=False), stg.jacobian(x, new_x=False))
        assert np.array_equal(reg.hessian(x, lam, 0.7, new_x=False),
                              stg.hessian(x, lam, 0.7, new_x=False))
        x += 0.01 * rng.standard_normal(it.num_x)        # same address, new values
        lam *= 1.1
    assert reg.num_registered_uploads == 3 and stg.num_registered_uploads == 0
    assert len(reg._registered) == 2                     # x and lam, registered once each
    reg.close()
    assert not reg._registered
    # the sampled compare works without this object's own copy of x
    smp = NlpCallbacks(it, "cyipopt", x_check="sampled")
    j0 = smp.jacobian(x).copy()
    up = smp.num_x_uploads
    smp.objective(x)
    assert smp.num_x_uploads == up and smp.num_registered_uploads == 1
    x[::max(1, x.size // 4096)] += 1e-3
    assert not np.array_equal(smp.jacobian(x), j0) and smp.num_x_uploads == up + 1
    # a non-contiguous / converted input is a temporary: staged, never registered
    n0 = len(smp._registered)
    smp.objective(x.astype(np.float32))
    assert len(smp._registered) == n0
    smp.close()


@pytest.mark.parametrize("name,K,nodes,kw", [
    ("cart_pole_swing_up", 700, 4, {}),
    ("double_pendulum", 6, [4, 7, 2, 10, 3, 5], dict(max_tile_nodes=20)),
    ("multiphase_sliding_mass", 60, [3, 5, 4] * 20, dict(max_tile_nodes=24)),
    ("delta_iii_launch_vehicle", 30, 4, {}),
])
def test_no_write_outside_the_output_arrays(name, K, nodes, kw):
    """Guard bands around every input and output (compute-sanitizer is not available
    on this pool): one arena, 512-double sentinels between the arrays, every
    callback evaluated, every sentinel and every input intact afterwards."""
    import torch
    low, _, scal = build_case(getattr(examples, name)(), "lobatto", K, nodes, oracle=False, **kw)
    S = low.S
    eng = make_engine(low, scal)
    sizes = dict(x=S.num_x, lam=S.num_c, sigma=1, f=1, grad=S.num_x, c=S.num_c, dy=S.num_dy,
                 jac=S.nnz_g, hess=S.nnz_h)
    G = 512
    total = sum(sizes.values()) + G * (len(sizes) + 1)
    SENT = -7.25e77
    arena = torch.full((total,), SENT, dtype=torch.float64, device="cuda")
    views, off = {}, G
    for k, n in sizes.items():
        views[k] = arena[off:off + n]
        off += n + G
    rng = np.random.default_rng(9)
    lo, hi = (0.3, 0.45) if name == "delta_iii_launch_vehicle" else (-0.5, 0.5)
    views["x"].copy_(torch.from_numpy(rng.uniform(lo, hi, S.num_x)))
    views["lam"].copy_(torch.from_numpy(rng.standard_normal(S.num_c)))
    views["sigma"].fill_(0.8)
    x0, l0 = views["x"].clone(), views["lam"].clone()
    what = E.EVAL_C | E.EVAL_DY | E.EVAL_JAC | E.EVAL_HESS | E.EVAL_F | E.EVAL_GRAD
    for w in (what, E.EVAL_JAC | E.EVAL_HESS, E.EVAL_JAC, E.EVAL_HESS, E.EVAL_C | E.EVAL_DY):
        eng.eval_ptr(w, views["x"], lam=views["lam"], sigma=views["sigma"],
                     f=views["f"], grad=views["grad"], c=views["c"], dy=views["dy"],
                     jac=views["jac"], hess=views["hess"])
    torch.cuda.synchronize()
    off = 0
    for k, n in sizes.items():                     # guard band in front of every array
        assert bool((arena[off:off + G] == SENT).all()), f"write in front of {k}"
        off += G + n
    assert bool((arena[off:off + G] == SENT).all()), "write behind the last array"
    assert torch.equal(views["x"], x0) and torch.equal(views["lam"], l0)
    for k in ("f", "grad", "c", "dy", "jac", "hess"):   # and every slot was written
        assert not bool((views[k] == SENT).any()), f"unwritten slot in {k}"
        assert bool(torch.isfinite(views[k]).all()), k


@pytest.mark.parametrize("name,K", [("delta_iii_launch_vehicle", 12), ("multiphase_sliding_mass", 9)])
def test_shared_expression_bodies_change_no_bit(name, K):
    """Phases that differ in literals only run ONE instantiation of the tile function
    (``codegen.share_groups``): same operations on the same doubles, so every output
    equals the one-instantiation-per-phase kernel's bit for bit."""
    low_s, _, scal = build_case(getattr(examples, name)(), "lobatto", K, 4, oracle=False)
    low_u, _, _ = build_case(getattr(examples, name)(), "lobatto", K, 4, oracle=False,
                             share_bodies=False)
    assert any(lay.leader != q for q, lay in enumerate(low_s.layouts))
    assert all(lay.leader == q and not lay.kc for q, lay in enumerate(low_u.layouts))
    rng = np.random.default_rng(11)
    x = rng.uniform(0.1, 0.4, low_s.S.num_x)
    lam = rng.standard_normal(low_s.S.num_c)
    a = make_engine(low_s, scal).eval_host(ALL, x, lam, 0.6)
    b = make_engine(low_u, scal).eval_host(ALL, x, lam, 0.6)
    for k in b:
        assert np.array_equal(a[k], b[k]), k

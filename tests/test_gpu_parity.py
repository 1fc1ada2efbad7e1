"""GPU parity: CUDA engine (through the C ABI) vs the CPU oracle, same inputs.

Tolerance is the one BASELINE.json's north_star states for fp64 values:
1e-12 relative (measured against max(|ref|, 1e-2*scale) so entries that are
pure cancellation noise are judged in units of the vector's scale); sparsity
indices are compared bit-exactly in the CPU tests (tests/test_structure.py).
"""
import numpy as np
import pytest

from helpers import (RAGGED_NODES, RAGGED_SIZES, build_case, make_engine, max_err)
from pycollo_b200 import engine as E
from examples import problems as examples

pytestmark = pytest.mark.gpu
RTOL = 1e-12
ALL = E.EVAL_C | E.EVAL_DY | E.EVAL_JAC | E.EVAL_HESS | E.EVAL_F | E.EVAL_GRAD

CASES = [
    ("brachistochrone", "lobatto", 10, 4, None, {}),
    ("brachistochrone", "radau", 10, 4, None, {}),
    ("cart_pole_swing_up", "lobatto", 10, 4, None, {}),
    ("cart_pole_swing_up", "radau", 40, 4, None, dict(max_tile_nodes=16)),
    ("hypersensitive", "lobatto", 6, RAGGED_NODES, RAGGED_SIZES, dict(max_tile_nodes=20)),
    ("hypersensitive", "radau", 6, RAGGED_NODES, RAGGED_SIZES, {}),
    ("double_pendulum", "lobatto", 6, RAGGED_NODES, RAGGED_SIZES, dict(max_tile_nodes=20)),
    ("double_pendulum", "radau", 10, 4, None, {}),
    ("cart_pole_swing_up", "lobatto", 2000, 4, None, {}),
    ("free_flying_robot", "lobatto", 40, [4, 6] * 20, None, {}),
    ("free_flying_robot", "radau", 300, 4, None, {}),
    ("space_shuttle_reentry", "radau", 6, RAGGED_NODES, RAGGED_SIZES, dict(max_tile_nodes=20)),
    ("space_shuttle_reentry", "lobatto", 400, 5, None, {}),
    ("multiphase_sliding_mass", "lobatto", 7, 4, None, {}),
    ("multiphase_sliding_mass", "radau", 200, 3, None, {}),
    ("delta_iii_launch_vehicle", "lobatto", 4, 4, None, {}),
    ("delta_iii_launch_vehicle", "lobatto", 500, 4, None, {}),
]


def _check(out, B, x, lam, sigma):
    errs = dict(
        f=max_err(out["f"], [B.J(x)]), grad=max_err(out["grad"][0], B.g(x)),
        c=max_err(out["c"][0], B.c(x)), dy=max_err(out["dy"][0], B.dy(x)),
        jac=max_err(out["jac"][0], B.G_nonzeros(x)),
        hess=max_err(out["hess"][0], B.H_nonzeros(x, sigma, lam)))
    bad = {k: v for k, v in errs.items() if not v <= RTOL}
    assert not bad, f"parity violated: {bad} (all: {errs})"


@pytest.mark.parametrize("name,method,K,nodes,sizes,kw", CASES)
def test_all_callbacks_match_oracle(cuda_device, name, method, K, nodes, sizes, kw):
    ocp = getattr(examples, name)()
    low, B, scal = build_case(ocp, method, K, nodes, sizes, seed=1, **kw)
    eng = make_engine(low, scal)
    rng = np.random.default_rng(7)
    for _ in range(2):
        x = rng.uniform(-0.5, 0.5, low.S.num_x)
        if name == "delta_iii_launch_vehicle":
            # keep the vehicle outside the Earth: exp(-(|r| - R_E)/h_0) overflows
            # for random positions near the centre (in the oracle just the same)
            x = rng.uniform(0.3, 0.45, low.S.num_x)
        lam = rng.standard_normal(low.S.num_c)
        sigma = float(rng.uniform(0.2, 2.0))
        out = eng.eval_host(ALL, x, lam, sigma)
        _check(out, B, x, lam, sigma)
        # single-output variants are separate NVRTC compilations (different FMA
        # contraction), so they agree with the fused launch to rounding only
        assert max_err(eng.eval_host(E.EVAL_JAC, x)["jac"], out["jac"]) <= 1e-13
        assert max_err(eng.eval_host(E.EVAL_HESS, x, lam, sigma)["hess"], out["hess"]) <= 1e-13
        assert max_err(eng.eval_host(E.EVAL_C, x)["c"], out["c"]) <= 1e-13
        jh = eng.eval_host(E.EVAL_JAC | E.EVAL_HESS, x, lam, sigma)
        assert max_err(jh["jac"], out["jac"]) <= 1e-13
        assert max_err(jh["hess"], out["hess"]) <= 1e-13


def test_golden_brachistochrone_pins(cuda_device):
    """tests/unit/test_iteration.py:305-385 of the reference, through the CUDA path."""
    from helpers import GOLDEN
    ocp = examples.brachistochrone()
    ocp.initialise()
    backend = ocp._backend
    it = backend.mesh_iterations[0]
    it.generate_nlp()
    g = np.load(f"{GOLDEN}/iteration_scaling_brachistochrone.npz")
    x = g["x_tilde"]
    assert it.num_x == 125 and it.num_c == 90
    np.testing.assert_almost_equal(backend.evaluate_J(x), 0.8243386694458454)
    expect_g = np.zeros(125)
    expect_g[124] = 10
    np.testing.assert_allclose(backend.evaluate_g(x), expect_g)
    np.testing.assert_allclose(backend.evaluate_c(x), np.zeros(90), atol=10e-2)
    rows, cols = backend.evaluate_G_structure()
    assert backend.evaluate_G_num_nonzero() == len(rows) == len(backend.evaluate_G_nonzeros(x))
    assert np.all(np.diff(cols) >= 0)                  # CCS: column-major


def test_golden_double_pendulum_pins(cuda_device):
    """tests/unit/test_iteration.py:290-336 of the reference."""
    from helpers import GOLDEN
    ocp = examples.double_pendulum()
    ocp.initialise()
    backend = ocp._backend
    it = backend.mesh_iterations[0]
    it.generate_nlp()
    g = np.load(f"{GOLDEN}/iteration_scaling_double_pendulum.npz")
    assert it.num_x == 190 and it.num_c == 121
    assert backend.evaluate_J(g["x_tilde"]) == 100
    expect_g = np.zeros(190)
    expect_g[186] = 1000
    np.testing.assert_allclose(backend.evaluate_g(g["x_tilde"]), expect_g)


def test_batched_instances(cuda_device):
    ocp = examples.cart_pole_swing_up()
    low, B, scal = build_case(ocp, "lobatto", 10, 4, seed=3)
    nb = 16
    eng = make_engine(low, scal, batch=nb)
    rng = np.random.default_rng(11)
    X = rng.uniform(-0.5, 0.5, (nb, low.S.num_x))
    L = rng.standard_normal((nb, low.S.num_c))
    sig = rng.uniform(0.5, 1.5, nb)
    out = eng.eval_host(ALL, X, L, sig)
    for i in (0, 5, nb - 1):
        one = {k: v[i:i + 1] for k, v in out.items()}
        _check(one, B, X[i], L[i], float(sig[i]))


def test_device_pointers_and_determinism(cuda_device):
    import torch
    ocp = examples.cart_pole_swing_up()
    low, B, scal = build_case(ocp, "lobatto", 500, 4, seed=5)
    eng = make_engine(low, scal)
    S = low.S
    rng = np.random.default_rng(2)
    x = rng.uniform(-0.5, 0.5, S.num_x)
    lam = rng.standard_normal(S.num_c)
    dx = torch.from_numpy(x).cuda()
    dl = torch.from_numpy(lam).cuda()
    jac = torch.full((S.nnz_g,), float("nan"), dtype=torch.float64, device="cuda")
    hes = torch.full((S.nnz_h,), float("nan"), dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    eng.eval_ptr(E.EVAL_JAC | E.EVAL_HESS, dx, lam=dl, jac=jac, hess=hes, stream=stream)
    torch.cuda.synchronize()
    j1, h1 = jac.cpu().numpy().copy(), hes.cpu().numpy().copy()
    assert max_err(j1, B.G_nonzeros(x)) <= RTOL
    assert max_err(h1, B.H_nonzeros(x, 1.0, lam)) <= RTOL
    for _ in range(3):                                    # bitwise reproducible
        jac.fill_(float("nan"))
        hes.fill_(float("nan"))
        eng.eval_ptr(E.EVAL_JAC | E.EVAL_HESS, dx, lam=dl, jac=jac, hess=hes, stream=stream)
        torch.cuda.synchronize()
        assert np.array_equal(jac.cpu().numpy(), j1) and np.array_equal(hes.cpu().numpy(), h1)


def test_const_resident_host_evaluation(cuda_device):
    """``PCX_EVAL_CONST_RESIDENT``: a host array that persists between evaluations gets
    its iterate-independent Jacobian slots once; later evaluations copy the other runs
    only -- and the array must still equal a full evaluation, also after a change of
    scaling and when another array is passed."""
    low, B, scal = build_case(examples.cart_pole_swing_up(), "lobatto", 4000, 4, seed=5)
    S = low.S
    ranges = S.G_constant_ranges()
    n_const = int((ranges[:, 1] - ranges[:, 0]).sum())
    assert n_const > 0.15 * S.nnz_g                      # the q1 / q1d columns
    eng = make_engine(low, scal)
    rng = np.random.default_rng(1)
    what = E.EVAL_JAC | E.EVAL_HESS
    jac, hes = np.full(S.nnz_g, np.nan), np.full(S.nnz_h, np.nan)
    other = np.full(S.nnz_g, np.nan)

    def run(flags, x, lam, out_j):
        eng._check(eng.lib.pcx_eval(eng.h, flags, E._ptr(x), E._ptr(lam), None, None, None, None,
                                    None, E._ptr(out_j), E._ptr(hes), E.PCX_HOST, None), "pcx_eval")
        return eng.last_d2h_bytes

    for k in range(4):
        x, lam = rng.uniform(-0.5, 0.5, S.num_x), rng.standard_normal(S.num_c)
        nbytes = run(what | E.EVAL_CONST_RESIDENT, x, lam, jac)
        full = 8 * (S.nnz_g + S.nnz_h)
        assert nbytes == (full if k == 0 else full - 8 * n_const)
        ref = eng.eval_host(what, x, lam, 1.0)
        assert np.array_equal(jac, ref["jac"][0]) and np.array_equal(hes, ref["hess"][0])
        assert max_err(jac, B.G_nonzeros(x)) <= RTOL
    assert run(what | E.EVAL_CONST_RESIDENT, x, lam, other) == full      # another array: everything
    assert np.array_equal(other, jac)
    eng.set_scaling(scal[0], scal[1], 2.0 * scal[2], scal[3])            # new scaling: everything
    assert run(what | E.EVAL_CONST_RESIDENT, x, lam, other) == full
    assert np.array_equal(other, eng.eval_host(E.EVAL_JAC, x)["jac"][0])
    assert run(E.EVAL_JAC | E.EVAL_CONST_RESIDENT, x, lam, other) == 8 * (S.nnz_g - n_const)


def test_pipelined_host_sweep(cuda_device):
    """``pcx_sweep_host``: pipelined host-space evaluations (upload / kernel / download
    of consecutive iterates overlap) leave in every host set exactly what one blocking
    ``pcx_eval`` per iterate leaves, with and without PCX_EVAL_CONST_RESIDENT."""
    import torch
    low, _, scal = build_case(examples.cart_pole_swing_up(), "lobatto", 5000, 4, seed=5, oracle=False)
    eng = make_engine(low, scal)
    S = low.S
    what = E.EVAL_JAC | E.EVAL_HESS
    rng = np.random.default_rng(4)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    n_const = int(np.diff(S.G_constant_ranges(), axis=1).sum())
    for flags in (what, what | E.EVAL_CONST_RESIDENT):
        sets = [dict(x=pin(rng.uniform(-0.5, 0.5, S.num_x)), lam=pin(rng.standard_normal(S.num_c)),
                     sigma=pin(np.array([0.5 + k])), jac=pin(np.full(S.nnz_g, np.nan)),
                     hess=pin(np.full(S.nnz_h, np.nan))) for k in range(3)]
        args = eng.make_args(sets)
        for count in (7, 3):                          # second round: constants already resident
            eng.sweep_host(flags, args, count)
            if flags & E.EVAL_CONST_RESIDENT:
                full = 8 * (S.nnz_g + S.nnz_h)
                assert eng.last_d2h_bytes == (full - 8 * n_const if n_const else full)
            for s_ in sets:
                # (the blocking call runs the Jacobian-only and Hessian-only variants, the
                # sweep the fused one: separate compilations, rounding-level differences)
                ref = eng.eval_host(what, s_["x"].numpy(), s_["lam"].numpy(), float(s_["sigma"][0]))
                assert max_err(s_["jac"].numpy(), ref["jac"][0]) <= 1e-13
                assert max_err(s_["hess"].numpy(), ref["hess"][0]) <= 1e-13
    with pytest.raises(E.PcxError, match=">= 2 host argument sets"):
        eng.sweep_host(what, eng.make_args(sets[:1]), 2)


def test_independent_sweep_is_bitwise_the_ordered_result(cuda_device):
    """``pcx_eval_many(PCX_EVAL_INDEPENDENT)``: evaluations declared independent are not
    ordered against each other on the device (no dependency wait, rotating scratch
    sets); every one of them must still produce exactly what stream order produces."""
    import torch
    for name, K in (("cart_pole_swing_up", 6000), ("multiphase_sliding_mass", 900)):
        low, _, scal = build_case(getattr(examples, name)(), "lobatto", K, 4, seed=5, oracle=False)
        eng = make_engine(low, scal)
        S = low.S
        what = E.EVAL_JAC | E.EVAL_HESS | E.EVAL_C | E.EVAL_F | E.EVAL_GRAD
        g = torch.Generator(device="cuda").manual_seed(3)
        z = lambda n: torch.full((n,), float("nan"), dtype=torch.float64, device="cuda")
        sets = [dict(x=torch.rand(S.num_x, dtype=torch.float64, device="cuda", generator=g) - 0.5,
                     lam=torch.randn(S.num_c, dtype=torch.float64, device="cuda", generator=g),
                     f=z(1), grad=z(S.num_x), c=z(S.num_c), jac=z(S.nnz_g), hess=z(S.nnz_h))
                for _ in range(5)]
        args = eng.make_args(sets)
        st = torch.cuda.current_stream().cuda_stream
        eng.eval_many(what, args, 5, stream=st, gate=False, timed=False)
        torch.cuda.synchronize()
        ref = [{k: s[k].clone() for k in ("f", "grad", "c", "jac", "hess")} for s in sets]
        for s in sets:
            for k in ("f", "grad", "c", "jac", "hess"):
                s[k].fill_(float("nan"))
        eng.eval_many(what | E.EVAL_INDEPENDENT, args, 35, stream=st, gate=True, timed=True)
        torch.cuda.synchronize()
        for s, r in zip(sets, ref):
            for k, v in r.items():
                assert torch.equal(s[k], v), (name, k)
    with pytest.raises(E.PcxError, match="distinct argument sets"):
        eng.eval_many(what | E.EVAL_INDEPENDENT, eng.make_args(sets[:2]), 4, stream=st)


@pytest.mark.parametrize("method", ["lobatto", "radau"])
def test_full_size_config2_properties(method, cuda_device):
    """BASELINE config 2 (cart-pole, 10^5 nodes) under both schemes -- Radau is the
    one the config names: oracle comparison at full size plus size-independent
    properties (linearity of H in (sigma, lam)).  The patterns behind the counts are
    pinned to the reference at N = 31 (tests/test_reference_goldens.py: cart_pole_lobatto
    1172 / 155, cart_pole_radau 941 / 150 = the same per-section formulas)."""
    ocp = examples.cart_pole_swing_up()
    low, B, scal = build_case(ocp, method, 33333, 4, seed=9)
    S = low.S
    # Lobatto: 38(N-1) + N + 1 and 5N.  Radau: the exactly-zero
    # last integration column / last weight of every section prune 23 Jacobian entries per
    # section, and the phase's last node (zero weight everywhere) carries no Hessian block
    want = (3899963, 500000) if method == "lobatto" else (3133303, 499995)
    assert (S.num_x, S.num_c) == (500001, 399997)
    assert (S.nnz_g, S.nnz_h) == want, (S.nnz_g, S.nnz_h)
    eng = make_engine(low, scal)
    rng = np.random.default_rng(0)
    x = rng.uniform(-0.5, 0.5, S.num_x)
    lam = rng.standard_normal(S.num_c)
    out = eng.eval_host(ALL, x, lam, 1.0)
    _check(out, B, x, lam, 1.0)
    h2 = eng.eval_host(E.EVAL_HESS, x, 3.0 * lam, 3.0)["hess"][0]
    assert max_err(h2, 3.0 * out["hess"][0]) <= 1e-13
    h0 = eng.eval_host(E.EVAL_HESS, x, 0.0 * lam, 0.0)["hess"][0]
    assert np.all(h0 == 0.0)


def test_iteration_scaling_from_sparse_jacobian(cuda_device):
    """N1 (scaling.py:346-430): W from row norms of the sparse G at the guess must
    equal the reference formula applied to the oracle's (dense-able, small) G."""
    from oracle.blockwise import BlockwiseNLP
    from helpers import oracle_meshes
    ocp = examples.free_flying_robot()
    ocp.initialise()
    backend = ocp._backend
    it = backend.mesh_iterations[0]
    it.generate_nlp()
    B = BlockwiseNLP(ocp, backend.ir.full_bounds, oracle_meshes(it.mesh.p))
    rows, cols = B.G_structure()
    G = np.zeros((B.num_c, B.num_x))
    G[rows, cols] = B.G_nonzeros(it.guess_x_tilde)
    norm = np.sqrt((G ** 2).sum(axis=1))
    ph = backend.ir.phases[0]
    N = it.mesh.N[0]
    W = it.scaling.W_ocp
    np.testing.assert_allclose(W[:6], 1.0 / it.scaling.V_ocp[:6], rtol=1e-14)
    p0 = 6 * (N - 1)
    expect_path = 1.0 / norm[p0:p0 + 2 * N].reshape(2, N).mean(axis=1)
    np.testing.assert_allclose(W[6:8], expect_path, rtol=1e-12)
    np.testing.assert_allclose(W[8], 1.0 / it.scaling.V_ocp[10], rtol=1e-14)
    # callbacks now run with that scaling
    c = backend.evaluate_c(it.guess_x_tilde)
    B2 = BlockwiseNLP(ocp, backend.ir.full_bounds, oracle_meshes(it.mesh.p), W_ocp=W, w=1.0)
    assert max_err(c, B2.c(it.guess_x_tilde)) <= RTOL
    cb = backend.nlp_callbacks()
    lam = np.random.default_rng(0).standard_normal(it.num_c)
    hr, hc = cb.hessianstructure()
    H = np.zeros((it.num_x, it.num_x))
    H[hr, hc] = cb.hessian(it.guess_x_tilde, lam, 0.5)
    br, bc = B2.H_structure()
    Hb = np.zeros_like(H)
    Hb[bc, br] = B2.H_nonzeros(it.guess_x_tilde, 0.5, lam)      # lower triangle
    assert max_err(H, Hb) <= RTOL

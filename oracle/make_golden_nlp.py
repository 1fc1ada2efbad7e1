"""Golden NLP vectors produced by EXECUTING THE UNMODIFIED REFERENCE PACKAGE.

TEST INFRASTRUCTURE ONLY -- run in the build container where ``/root/reference``
is mounted; the GPU box never sees the reference, only the committed
``tests/golden/nlp_*.npz`` files this script writes.

How: ``/root/reference/pycollo`` is imported as it is.  Its two absent
third-party dependencies are replaced by stand-ins under ``oracle/refshim``:
``pyproprop`` (validated descriptors / option sets) and ``casadi`` (a
sympy-backed ``SX``/``DM``/``Function``/``nlpsol``; see its module docstring for
the modelling rules and their limits).  The same user scripts the CUDA tests
use (``examples/problems.py``) are built on the reference's own
``pycollo.OptimalControlProblem`` and ``ocp.initialise()`` runs end to end:
``Casadi.generate_nlp_function_callables`` (``backend.py:1403-1679``),
``IterationScaling`` (``scaling.py:124-454``), ``Iteration.generate_bounds``
(``iteration.py:396-453``), guess interpolation (``iteration.py:86-194``),
``create_nlp_solver`` (``backend.py:1681-1693``).  Then the reference's own
``evaluate_*`` members (``backend.py:1713-1771``) and the nlpsol oracle function
``nlp_hess_l`` are evaluated at the scaled guess and at seeded random iterates.

The reference's unit tests for this path (``tests/unit/test_iteration.py``,
``test_iteration_scaling.py``, ``test_scaling.py``, ``test_quadrature.py``,
``test_initialisation.py``, ``test_optimal_control_problem.py``) pass under these
stand-ins (48 passed; the 4 failures of ``test_utils.py`` compare CasADi's
expression *printer* output), which is what validates the stand-ins themselves.

File schema (per problem/scheme; names follow the reference's JSON dump
``iteration.py:1216-1234`` where one exists):
  num_x num_c G_nnz H_nnz  G_row G_col  H_row H_col   (CCS order; H upper triangle)
  x[k] J[k] g[k] c[k] dy[k] G_data[k] lam[k] sigma[k] H_data[k]   k = 0..n_pts-1
  J_dbl g_dbl c_dbl dy_dbl G_data_dbl H_data_dbl  (same, plain double evaluation)
  J_err g_err c_err dy_err G_data_err H_data_err  (per element: first-order running-error
      bound of evaluating the reference's own expression in fp64 -- inputs exact, unit
      roundoff 2^-53 per operation; large where x = V*x_tilde + r cancels)
  V r V_ocp r_ocp w W_ocp W  guess_x x_bnd_l x_bnd_u c_bnd_l c_bnd_u
  per phase p: tau_p N_K_p h_K_p sI_{data,indices,indptr}_p sA_{...}_p Wq_p
Values are evaluated with mpmath at 40 digits and rounded once (``casadi.PRECISE``);
``double_vs_exact`` records how far plain double evaluation of the same
expressions is from them.

Usage:  python oracle/make_golden_nlp.py [name ...]
"""
import contextlib
import io
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference():
    """``import pycollo`` from the reference tree with the stand-ins in place."""
    for p in (ROOT, os.path.join(HERE, "refshim"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:                       # pycollo/vis/plot.py names it at import
            mp = types.ModuleType("matplotlib")
            mp.pyplot = types.ModuleType("matplotlib.pyplot")
            sys.modules["matplotlib"] = mp
            sys.modules["matplotlib.pyplot"] = mp.pyplot
    import casadi
    assert casadi.__version__ == "refshim-sympy"
    import pycollo
    assert os.path.realpath(pycollo.__file__).startswith(REF)
    return pycollo, casadi


def csr_parts(prefix, m):
    m = m.tocsr()
    return {f"{prefix}_data": m.data.astype(np.float64),
            f"{prefix}_indices": m.indices.astype(np.int64),
            f"{prefix}_indptr": m.indptr.astype(np.int64)}


def _dense_bound(casadi, fn, index, n, *vals):
    """Error bounds of a (possibly sparse) column output as a dense vector."""
    out = np.zeros(n)
    keys = fn.sx_out()[index]._keys_ccs()
    if keys:
        out[[k[0] for k in keys]] = casadi.error_bounds(fn, *vals)[index]
    return out


def generate(name, pycollo, casadi, n_random=2):
    from examples.cases import build_golden_problem
    ocp = build_golden_problem(name, api=pycollo)
    ocp.settings.display_mesh_result_graph = False
    t0 = time.time()
    casadi.PRECISE = True
    with contextlib.redirect_stdout(io.StringIO()):
        ocp.initialise()                                  # optimal_control_problem.py:316-337
    backend = ocp._backend
    it = backend.mesh_iterations[0]
    out = {"num_x": np.int64(it.num_x), "num_c": np.int64(it.num_c),
           "num_phases": np.int64(len(backend.p))}
    # ---- what the iteration derived (N1 scaling, N3 guess + bounds) ----------
    sc = it.scaling
    out.update(V=sc.V, r=sc.r, V_ocp=sc.V_ocp, r_ocp=sc.r_ocp, w=np.float64(sc.w),
               W_ocp=np.asarray(sc.W_ocp, dtype=np.float64), W=np.asarray(sc.W, dtype=np.float64),
               guess_x=np.asarray(it.guess_x, dtype=np.float64),
               x_bnd_l=it.x_bnd_l, x_bnd_u=it.x_bnd_u, c_bnd_l=it.c_bnd_l, c_bnd_u=it.c_bnd_u)
    m = it.mesh
    for ip in range(len(backend.p)):
        out[f"tau_{ip}"] = np.asarray(m.tau[ip], dtype=np.float64)
        out[f"N_K_{ip}"] = np.asarray(m.N_K[ip], dtype=np.int64)
        out[f"h_K_{ip}"] = np.asarray(m.h_K[ip], dtype=np.float64)
        out[f"sizes_{ip}"] = np.asarray(ocp.phases[ip].mesh.mesh_section_sizes, dtype=np.float64)
        out[f"Wq_{ip}"] = np.asarray(m.W_matrix[ip], dtype=np.float64)
        out.update(csr_parts(f"sI_{ip}", m.sI_matrix[ip]))
        out.update(csr_parts(f"sA_{ip}", m.sA_matrix[ip]))
    # ---- structure: the reference's own reader (backend.py:1747-1771) ------------
    G_row, G_col = backend.evaluate_G_structure()
    out.update(G_row=np.asarray(G_row, dtype=np.int64), G_col=np.asarray(G_col, dtype=np.int64),
               G_nnz=np.int64(backend.evaluate_G_num_nonzero()))
    jac_fn = backend.nlp_solver.get_function("nlp_jac_g")
    dead = casadi.jacobian_check(jac_fn.sx_out()[1])
    assert not dead, f"{name}: {len(dead)} G entries are structurally present but identically zero"
    hess_fn = backend.nlp_solver.get_function("nlp_hess_l")
    H = hess_fn.sx_out()[0]
    assert not casadi.jacobian_check(H), f"{name}: identically-zero H entries"
    hk = H._keys_ccs()
    out.update(H_row=np.array([k[0] for k in hk], dtype=np.int64),
               H_col=np.array([k[1] for k in hk], dtype=np.int64), H_nnz=np.int64(len(hk)))
    assert np.all(out["H_row"] <= out["H_col"])
    # ---- values at the scaled guess and at seeded random iterates -----------------
    rng = np.random.default_rng(sum(map(ord, name)))
    xs = [np.asarray(it.guess_x, dtype=np.float64)]
    xs += [rng.uniform(-0.5, 0.5, it.num_x) for _ in range(n_random)]
    # Delta III: the initial guess is a singular point of the dynamics (relative
    # wind speed exactly 0 -> d sqrt(.) = 1/0 in fp64; the reference's own Jacobian
    # is not finite there), so values are pinned at the random iterates only and
    # the W the reference derives at the guess is not a meaningful pin
    singular_guess = name.startswith("delta_iii")
    if singular_guess:
        xs = xs[1:] + [rng.uniform(-0.5, 0.5, it.num_x)]
    out["singular_guess"] = np.bool_(singular_guess)
    lams = [rng.standard_normal(it.num_c) for _ in xs]
    sigmas = [1.0, 0.7, -1.3][:len(xs)]
    keys = ("x", "J", "g", "c", "dy", "G_data", "lam", "sigma", "H_data")
    dkeys = ("J", "g", "c", "dy", "G_data", "H_data")
    acc = {k: [] for k in keys}
    dbl = {k: [] for k in dkeys}
    cond = {k: [] for k in dkeys}
    # the same expressions evaluated in plain double precision (what the reference
    # itself computes with, up to operation order): recorded beside the exact values
    # so a test can tell fp64 cancellation noise (x = V*x_tilde + r loses eps*|r|)
    # from a real discrepancy
    solver = backend.nlp_solver
    casadi.PRECISE = False
    d_fn = {nm: casadi.Function(nm + "_dbl", solver.get_function(nm).sx_in(),
                                solver.get_function(nm).sx_out())
            for nm in ("nlp_f", "nlp_grad_f", "nlp_g", "nlp_jac_g", "nlp_hess_l")}
    d_dy = casadi.Function("dy_dbl", [backend.x_var_iter], [backend.dy_iter])
    casadi.PRECISE = True
    dy_fn = casadi.Function("dy", [backend.x_var_iter], [backend.dy_iter])
    flat = lambda v: np.asarray(v, dtype=np.float64).reshape(-1)
    worst = 0.0
    for x, lam, sg in zip(xs, lams, sigmas):
        acc["x"].append(x)
        acc["J"].append(backend.evaluate_J(x))                         # backend.py:1713
        acc["g"].append(flat(backend.evaluate_g(x)))                   # :1717
        acc["c"].append(flat(backend.evaluate_c(x)))                   # :1722
        acc["dy"].append(flat(dy_fn(x)))                               # :1665
        acc["G_data"].append(flat(backend.evaluate_G_nonzeros(x)))     # :1738
        acc["lam"].append(lam)
        acc["sigma"].append(sg)
        acc["H_data"].append(flat(hess_fn(x, [], sg, lam).nonzeros()))
        dbl["J"].append(float(d_fn["nlp_f"](x, [])))
        dbl["g"].append(flat(d_fn["nlp_grad_f"](x, [])[1]))
        dbl["c"].append(flat(d_fn["nlp_g"](x, [])))
        dbl["dy"].append(flat(d_dy(x)))
        dbl["G_data"].append(flat(d_fn["nlp_jac_g"](x, [])[1].nonzeros()))
        dbl["H_data"].append(flat(d_fn["nlp_hess_l"](x, [], sg, lam).nonzeros()))
        # running-error bounds of the reference's own fp64 evaluation, per element
        eb = casadi.error_bounds
        cond["J"].append(_dense_bound(casadi, d_fn["nlp_f"], 0, 1, x, [])[0])
        cond["g"].append(_dense_bound(casadi, d_fn["nlp_grad_f"], 1, it.num_x, x, []))
        cond["c"].append(_dense_bound(casadi, d_fn["nlp_g"], 0, it.num_c, x, []))
        cond["dy"].append(_dense_bound(casadi, d_dy, 0, acc["dy"][-1].size, x))
        cond["G_data"].append(eb(d_fn["nlp_jac_g"], x, [])[1])
        cond["H_data"].append(eb(d_fn["nlp_hess_l"], x, [], sg, lam)[0])
        for k in ("G_data", "H_data"):
            a_, b_ = dbl[k][-1], acc[k][-1]
            if a_.size:
                floor = 1e-2 * max(1.0, float(np.max(np.abs(b_))))
                worst = max(worst, float(np.max(np.abs(a_ - b_) / np.maximum(np.abs(b_), floor))))
    for k in keys:
        out[k] = np.array(acc[k], dtype=np.float64)
    for k in dkeys:
        out[k + "_dbl"] = np.array(dbl[k], dtype=np.float64)
        out[k + "_err"] = np.array(cond[k], dtype=np.float64).reshape(out[k].shape)
    out["double_vs_exact"] = np.float64(worst)
    casadi.PRECISE = False
    np.savez_compressed(os.path.join(OUT, f"nlp_{name}.npz"), **out)
    print(f"{name}: num_x={it.num_x} num_c={it.num_c} nnz_G={int(out['G_nnz'])} "
          f"nnz_H={int(out['H_nnz'])} double-vs-exact {worst:.1e}  ({time.time() - t0:.0f} s)",
          flush=True)


def main(argv):
    os.makedirs(OUT, exist_ok=True)
    pycollo, casadi = import_reference()
    from examples.cases import GOLDEN_CASES
    for name in (argv or list(GOLDEN_CASES)):
        generate(name, pycollo, casadi)


if __name__ == "__main__":
    main(sys.argv[1:])

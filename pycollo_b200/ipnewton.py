"""Built-in host NLP solver: a primal-dual interior-point Newton method over the
cyipopt callback contract (``pycollo/nlp.py:36-76``).

The reference solves its NLP with IPOPT through ``ca.nlpsol`` (``pycollo/backend.py:
1681-1693, 1807-1827``).  IPOPT / cyipopt / CasADi are not installable in this image, so
``Cuda.solve_nlp`` falls back on this solver when cyipopt cannot be imported.  It is HOST
code by design -- so is IPOPT: the linear algebra of the Newton step (scipy's sparse LU
here, MUMPS there) is not on the path this package accelerates; every function value and
derivative it consumes comes from the CUDA callbacks.  It is not IPOPT: iteration counts
and solve times of the two are not comparable (``scipy.optimize``'s ``trust-constr`` was
tried first: on the brachistochrone it is still at a constraint violation of 6e-2 after
3000 iterations).
"""
import numpy as np
import scipy.sparse as sp


def solve(cb, x0, x_lo, x_hi, c_lo, c_hi, tol=1e-8, max_iter=500, verbose=False):
    """Primal-dual interior-point Newton method in the style of IPOPT's basic
    algorithm (Waechter & Biegler 2006, sections 2-3, with its filter acceptance test
    but without the restoration phase): slacks for inequality rows, log barrier for
    the variable bounds, exact Hessian of the Lagrangian, fraction-to-the-boundary
    rule, backtracking line search, diagonal regularisation when the reduced Hessian
    is not positive definite, monotone barrier update.  Deterministic: the sequence of
    iterates is a function of the callback values only.

    Returns ``SimpleNamespace(x, fun, nit, constr_violation, kkt_error, success, lam)``.
    """
    from types import SimpleNamespace
    import scipy.sparse.linalg as spla
    x0 = np.array(x0, dtype=float)
    n, m = len(x0), len(c_lo)
    gr, gc = cb.jacobianstructure()
    hr, hc = cb.hessianstructure()
    off = hr != hc
    ineq = np.flatnonzero(c_hi > c_lo)
    ns = len(ineq)
    lo = np.concatenate([x_lo, c_lo[ineq]])
    hi = np.concatenate([x_hi, c_hi[ineq]])
    # variables with coinciding bounds are parameters (IPOPT's make_parameter): the
    # Newton system is restricted to the free ones (P selects their columns)
    fixed = (hi - lo) <= 1e-13 * np.maximum(1.0, np.abs(lo))
    free_idx = np.flatnonzero(~fixed)
    has_lo = np.isfinite(lo) & (lo > -1e18) & ~fixed
    has_hi = np.isfinite(hi) & (hi < 1e18) & ~fixed
    # slack columns of the equality form  h(v) = c(x) - s (inequality rows) / - c_lo
    S_cols = sp.csr_matrix((-np.ones(ns), (ineq, np.arange(ns))), shape=(m, ns))
    target = np.where(c_hi > c_lo, 0.0, c_lo)

    def push(v):                       # strictly inside the bounds (IPOPT's kappa_1, kappa_2)
        v = v.copy()
        span = np.where(has_lo & has_hi, hi - lo, np.inf)
        pl = np.minimum(1e-2 * np.maximum(1.0, np.abs(lo)), 1e-2 * span)
        pu = np.minimum(1e-2 * np.maximum(1.0, np.abs(hi)), 1e-2 * span)
        v = np.where(has_lo, np.maximum(v, lo + pl), v)
        v = np.where(has_hi, np.minimum(v, hi - pu), v)
        return v

    def funcs(v):
        x = v[:n]
        c = np.array(cb.constraints(x), dtype=float)
        h = c - target
        h[ineq] -= v[n:]
        A = sp.hstack([sp.csr_matrix((np.array(cb.jacobian(x), dtype=float), (gr, gc)),
                                     shape=(m, n)), S_cols]).tocsr()
        g = np.concatenate([np.array(cb.gradient(x), dtype=float), np.zeros(ns)])
        return float(cb.objective(x)), g, h, A

    def lag_hess(v, lam):
        hv = np.array(cb.hessian(v[:n], lam, 1.0), dtype=float)
        H = sp.coo_matrix((np.concatenate([hv, hv[off]]),
                           (np.concatenate([hr, hc[off]]), np.concatenate([hc, hr[off]]))),
                          shape=(n, n)).tocsr()
        return sp.block_diag([H, sp.csr_matrix((ns, ns))]).tocsr() if ns else H

    def barrier(v, f, mu):
        val = f
        if has_lo.any():
            val -= mu * np.sum(np.log(v[has_lo] - lo[has_lo]))
        if has_hi.any():
            val -= mu * np.sum(np.log(hi[has_hi] - v[has_hi]))
        return val

    v = np.concatenate([x0, np.zeros(ns)])
    if ns:
        v[n:] = np.array(cb.constraints(x0), dtype=float)[ineq]
    v = push(v)
    v[fixed] = lo[fixed]
    P = sp.csr_matrix((np.ones(len(free_idx)), (free_idx, np.arange(len(free_idx)))),
                      shape=(n + ns, len(free_idx)))
    mu = 0.1
    zl = np.where(has_lo, mu / np.where(has_lo, v - lo, 1.0), 0.0)
    zu = np.where(has_hi, mu / np.where(has_hi, hi - v, 1.0), 0.0)
    lam = np.zeros(m)
    f, g, h, A = funcs(v)
    delta_last, it = 0.0, 0
    filt, filt_mu = [], None
    theta_init = float(np.sum(np.abs(h)))
    N = n + ns
    for it in range(1, max_iter + 1):
        dl = np.where(has_lo, v - lo, 1.0)
        du = np.where(has_hi, hi - v, 1.0)
        r_dual = (g + A.T @ lam - zl + zu) * ~fixed

        def err(mu_):
            comp = max(np.max(np.abs(zl * dl - mu_), initial=0.0, where=has_lo),
                       np.max(np.abs(zu * du - mu_), initial=0.0, where=has_hi))
            return max(np.max(np.abs(r_dual), initial=0.0), np.max(np.abs(h), initial=0.0), comp)

        e0 = err(0.0)
        if verbose:
            print(f"{it:4d} f={f:+.8e} |h|={np.max(np.abs(h), initial=0):.2e} err={e0:.2e} mu={mu:.1e}")
        if e0 <= tol:
            break
        while mu > tol / 10 and err(mu) <= 10 * mu:
            mu = max(tol / 10, min(0.2 * mu, mu ** 1.5))
        Sigma = np.where(has_lo, zl / dl, 0.0) + np.where(has_hi, zu / du, 0.0)
        W = lag_hess(v, lam)
        rhs_d = g + A.T @ lam - np.where(has_lo, mu / dl, 0.0) + np.where(has_hi, mu / du, 0.0)
        delta = 0.0
        Wf, Af, nf = P.T @ W @ P, A @ P, len(free_idx)
        for attempt in range(40):
            K = sp.bmat([[Wf + sp.diags((Sigma + delta)[free_idx]), Af.T],
                         [Af, -1e-9 * sp.identity(m)]], format="csc")
            try:
                sol = spla.splu(K).solve(-np.concatenate([rhs_d[free_idx], h]))
                dv, dlam = np.zeros(N), sol[nf:]
                dv[free_idx] = sol[:nf]
                curv = dv @ ((W @ dv) + (Sigma + delta) * dv)
                ok = np.all(np.isfinite(sol)) and curv > 1e-12 * (dv @ dv)
            except RuntimeError:
                ok = False
            if ok:
                break
            delta = max(1e-4, delta_last / 3) if delta == 0.0 else delta * 8
        delta_last = delta
        dzl = np.where(has_lo, mu / dl - zl - zl / dl * dv, 0.0)
        dzu = np.where(has_hi, mu / du - zu + zu / du * dv, 0.0)
        tau = max(0.99, 1.0 - mu)

        def max_step(val, dval, mask):
            neg = mask & (dval < 0)
            return min(1.0, np.min(-tau * val[neg] / dval[neg], initial=1.0))

        a_p = min(max_step(dl, dv, has_lo), max_step(du, -dv, has_hi))
        a_d = min(max_step(zl, dzl, has_lo), max_step(zu, dzu, has_hi))
        # globalisation: IPOPT's filter acceptance test (its eq. 18-20) with the
        # filter reset at every barrier update -- a trial point is taken if it
        # improves the constraint violation or the barrier objective sufficiently
        # and is not dominated by an earlier iterate of this barrier problem
        phi0 = barrier(v, f, mu)
        theta0 = np.sum(np.abs(h))
        dphi = rhs_d @ dv - (A.T @ lam) @ dv
        if mu != filt_mu:
            filt, filt_mu = [], mu
        alpha = a_p
        for ls in range(40):
            v_new = v + alpha * dv
            f_n, g_n, h_n, A_n = funcs(v_new)
            if np.isfinite(f_n) and np.all(np.isfinite(h_n)):
                th, ph = np.sum(np.abs(h_n)), barrier(v_new, f_n, mu)
                if theta0 <= 1e-4 * max(1.0, theta_init) and dphi < 0 and \
                        alpha * (-dphi) ** 2.3 > theta0 ** 1.1:
                    okay = ph <= phi0 + 1e-8 * alpha * dphi        # Armijo on the barrier
                else:
                    okay = th <= (1 - 1e-5) * theta0 or ph <= phi0 - 1e-5 * theta0
                if okay and all(th < (1 - 1e-5) * t_ or ph < p_ - 1e-5 * t_ for t_, p_ in filt):
                    break
            alpha *= 0.5
        filt.append((theta0, phi0))
        v, f, g, h, A = v_new, f_n, g_n, h_n, A_n
        lam = lam + alpha * dlam
        zl = np.where(has_lo, zl + a_d * dzl, 0.0)
        zu = np.where(has_hi, zu + a_d * dzu, 0.0)
        # keep the bound multipliers near the central path (IPOPT eq. 16)
        dl = np.where(has_lo, v - lo, 1.0)
        du = np.where(has_hi, hi - v, 1.0)
        zl = np.where(has_lo, np.clip(zl, mu / (1e10 * dl), 1e10 * mu / dl), 0.0)
        zu = np.where(has_hi, np.clip(zu, mu / (1e10 * du), 1e10 * mu / du), 0.0)
    x = v[:n]
    c = np.array(cb.constraints(x), dtype=float)
    viol = max(np.max(np.maximum(c_lo - c, 0), initial=0.0), np.max(np.maximum(c - c_hi, 0), initial=0.0))
    return SimpleNamespace(x=x, fun=float(cb.objective(x)), nit=it, constr_violation=float(viol),
                           kkt_error=float(e0), success=bool(e0 <= tol), lam=lam)

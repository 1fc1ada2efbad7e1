"""ORACLE -- mesh-refinement error of pycollo's Patterson-Rao scheme.  TEST ONLY.

CPU restatement (plain numpy loops) of
``pycollo/mesh_refinement.py:75-86``  (the p+1 "ph" mesh: same sections, one more
node each), ``:206-240`` (``phase_mesh_error``) and of the polynomial re-fit that
feeds it, ``pycollo/solution/solution_abc.py:60-142`` (per-section Legendre fit
of dy integrated from the section's first state value; plain polynomial fit of
the controls) + ``mesh_refinement.py:160-204`` (``construct_x_ph``).

Pinned: ``tests/golden/mesh_error_*.npz`` hold the outputs of the reference's own
``PattersonRaoMeshRefinement.phase_mesh_error`` (executed from
``/root/reference`` by ``oracle/make_golden.py`` on the reference's own ph
``Mesh``); ``tests/test_mesh_error.py`` checks this file against them.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
may import this module; the product never does.
"""
from __future__ import annotations

import numpy as np


def ph_section_nodes(N_K):
    """``mesh_refinement.py:80``: every section gets one more node."""
    return np.asarray(N_K, dtype=np.int64) + 1


def phase_mesh_error(dy_ph_vec, y_ph, sI_ph, stretch, N_K_ph, boundaries_ph):
    """``mesh_refinement.py:206-240`` for one phase.

    dy_ph_vec : state-major vector of f at every ph node (what ``dy_ph_callables``
                returns); y_ph : (n_y, N_ph); sI_ph : ph integration CSR.
    Returns (absolute (K, n_y, mmax), relative (K, n_y, mmax), max_relative (K,)).
    """
    n_y = y_ph.shape[0]
    K = len(N_K_ph)
    dy_ph = np.asarray(dy_ph_vec, dtype=float).reshape((-1, n_y), order="F")   # :208
    I_dy_ph = stretch * sI_ph.dot(dy_ph)                                       # :209
    mmax = int(max(N_K_ph)) - 1                                                # :212
    mesh_error = np.zeros((K, n_y, mmax))
    scale = np.zeros((K, n_y))
    for i_k, (i_start, m_k) in enumerate(zip(boundaries_ph[:-1],
                                             np.asarray(N_K_ph) - 1)):        # :216-224
        y_k = y_ph[:, i_start]
        Y_ph_k = (y_k + I_dy_ph[i_start:i_start + m_k]).T
        Y_k = y_ph[:, i_start + 1:i_start + 1 + m_k]
        mesh_error[i_k, :, :m_k] = Y_ph_k - Y_k
        scale[i_k, :] = np.max(np.abs(Y_k), axis=1) + 1
    absolute = np.abs(mesh_error)                                              # :226
    relative = np.zeros_like(absolute)
    for i_k in range(K):                                                       # :229-234
        for i_y in range(n_y):
            for i_m in range(int(N_K_ph[i_k]) - 1):
                relative[i_k, i_y, i_m] = absolute[i_k, i_y, i_m] / (1 + scale[i_k, i_y])
    max_rel = np.zeros(K)
    for i_k in range(K):                                                       # :237-240
        max_rel[i_k] = np.max(relative[i_k, :, :])
    return absolute, relative, max_rel


def fit_section_polys(tau, y, dy, u, T, boundaries, N_K, method="lobatto", period=2.0):
    """``solution_abc.py:60-142``: per state and section a Legendre fit of
    ``dy * T / period`` integrated from ``y[start]``; per control a polynomial fit."""
    K = len(N_K)
    y_polys = np.empty((y.shape[0], K), dtype=object)
    u_polys = np.empty((u.shape[0], K), dtype=object)
    sf = T / period
    for i_y in range(y.shape[0]):
        for i_k, (a, b) in enumerate(zip(boundaries[:-1], boundaries[1:])):
            t_k = tau[a:b + 1]
            dy_k = dy[i_y, a:b + 1] * sf
            if method == "lobatto":                                            # :86-91
                p = np.polynomial.Legendre.fit(t_k, dy_k, deg=int(N_K[i_k]) - 1, window=[0, 1])
            else:                                                              # :117-123
                p = np.polynomial.Legendre.fit(t_k[:-1], dy_k[:-1], deg=int(N_K[i_k]) - 2,
                                               domain=[t_k[0], t_k[-1]], window=[0, 1])
            y_polys[i_y, i_k] = p.integ(k=y[i_y, a])
    for i_u in range(u.shape[0]):
        for i_k, (a, b) in enumerate(zip(boundaries[:-1], boundaries[1:])):
            t_k = tau[a:b + 1]
            u_polys[i_u, i_k] = np.polynomial.Polynomial.fit(
                t_k, u[i_u, a:b + 1], deg=int(N_K[i_k]) - 1, window=[0, 1])
    return y_polys, u_polys


def interpolate_to_ph(vals, polys, boundaries, boundaries_ph, tau_ph):
    """``mesh_refinement.py:162-190``: copy the section-boundary values, evaluate
    the section polynomials at the interior ph nodes."""
    out = np.zeros((vals.shape[0], len(tau_ph)))
    out[:, boundaries_ph] = vals[:, boundaries]
    for i_var in range(vals.shape[0]):
        for i_k, (a, b) in enumerate(zip(boundaries_ph[:-1], boundaries_ph[1:])):
            out[i_var, a + 1:b] = polys[i_var, i_k](tau_ph[a + 1:b])
    return out

"""Solution post-processing that feeds the mesh-refinement error pass.

Mirrors, for the data the error evaluation needs, ``pycollo/solution/
casadi_solution.py:15-84`` (unscale x, slice per phase, state derivatives from
the backend's ``dy_iter_callable`` -- here one ``PCX_EVAL_DY`` launch) and
``pycollo/solution/solution_abc.py:60-142`` (per-section polynomial re-fit:
Legendre fit of ``dy * T/2`` integrated from the section's first state value;
plain polynomial fit of the controls).  Host-side plumbing (SURVEY.md section 8
row N2); all NLP function values come from the CUDA engine.
"""
from __future__ import annotations

import collections

import numpy as np

from . import engine as _engine

PhaseSolutionData = collections.namedtuple(
    "PhaseSolutionData", ("tau", "y", "dy", "u", "q", "t", "t0", "tF", "T",
                          "stretch", "shift", "time"))
Polys = collections.namedtuple("Polys", ("y", "dy", "u"))
_PERIOD = 2.0                      # pycollo/mesh.py: tau in [-1, 1]


class Solution:
    """``CasadiSolution`` for an iterate ``x_tilde`` of a mesh iteration."""

    def __init__(self, iteration, x_tilde, J=None):
        self.it = iteration
        self.backend = iteration.backend
        self.ocp = iteration.ocp
        self.tau = iteration.mesh.tau
        self.x = np.asarray(x_tilde, dtype=np.float64)
        self.J = J
        self.process_solution()

    def process_solution(self):
        self.extract_full_solution()
        self.set_user_attributes()
        self.interpolate_solution(self.ocp.settings.quadrature_method)

    def extract_full_solution(self):
        """``casadi_solution.py:15-31, 43-84``."""
        it = self.it
        x = it.scaling.unscale_x(self.x)
        dy = it.evaluate(_engine.EVAL_DY, self.x)["dy"][0]
        data = []
        for ip, (ph, t) in enumerate(zip(self.backend.ir.phases, it.S.ph)):
            N = t.N
            y = x[it.y_slices[ip]].reshape(ph.n_y, N) if ph.n_y else np.empty((0, N))
            dyp = dy[it.dy_slices[ip]].reshape(ph.n_y, N) if ph.n_y else np.empty((0, N))
            u = x[it.u_slices[ip]].reshape(ph.n_u, N) if ph.n_u else np.empty((0, N))
            q = x[it.q_slices[ip]]
            tt = x[it.t_slices[ip]]
            t0 = tt[0] if ph.t_needed[0] else float(ph.t0)
            tF = tt[-1] if ph.t_needed[1] else float(ph.tF)
            T = tF - t0
            stretch, shift = T / 2, (t0 + tF) / 2
            tau = np.asarray(self.tau[ip])
            data.append(PhaseSolutionData(tau, y, dyp, u, q, tt, t0, tF, T, stretch,
                                          shift, tau * stretch + shift))
        self.phase_data = tuple(data)
        self._s = x[it.s_slice]
        if self.J is not None:
            self.objective = self.J / it.scaling.w

    def set_user_attributes(self):
        pd = self.phase_data
        self.state = tuple(p.y for p in pd)
        self.state_derivative = tuple(p.dy for p in pd)
        self.control = tuple(p.u for p in pd)
        self.integral = tuple(p.q for p in pd)
        self.time = tuple(p.t for p in pd)
        self.parameter = self._s
        self.initial_time = tuple(p.t0 for p in pd)
        self.final_time = tuple(p.tF for p in pd)

    def interpolate_solution(self, method):
        """``solution_abc.py:60-142`` (Lobatto ``:60-101``, Radau ``:103-142``)."""
        self.phase_polys = []
        mesh = self.it.mesh
        for ip, p_data in enumerate(self.phase_data):
            K, N_K = mesh.K[ip], mesh.N_K[ip]
            bnd = mesh.mesh_index_boundaries[ip]
            ny, nu = p_data.y.shape[0], p_data.u.shape[0]
            y_polys = np.empty((ny, K), dtype=object)
            dy_polys = np.empty((ny, K), dtype=object)
            u_polys = np.empty((nu, K), dtype=object)
            sf = p_data.T / _PERIOD
            for i_k, (a, b) in enumerate(zip(bnd[:-1], bnd[1:])):
                t_k = p_data.tau[a:b + 1]
                n = int(N_K[i_k])
                for i_y in range(ny):
                    d_k = p_data.dy[i_y, a:b + 1]
                    if method == "lobatto":
                        dy_polys[i_y, i_k] = np.polynomial.Legendre.fit(
                            t_k, d_k, deg=n - 1, window=[0, 1])
                        scaled = np.polynomial.Legendre.fit(
                            t_k, d_k * sf, deg=n - 1, window=[0, 1])
                    else:
                        dom = [t_k[0], t_k[-1]]
                        dy_polys[i_y, i_k] = np.polynomial.Legendre.fit(
                            t_k[:-1], d_k[:-1], deg=n - 2, domain=dom, window=[0, 1])
                        scaled = np.polynomial.Legendre.fit(
                            t_k[:-1], d_k[:-1] * sf, deg=n - 2, domain=dom, window=[0, 1])
                    y_polys[i_y, i_k] = scaled.integ(k=p_data.y[i_y, a])
                for i_u in range(nu):
                    u_polys[i_u, i_k] = np.polynomial.Polynomial.fit(
                        t_k, p_data.u[i_u, a:b + 1], deg=n - 1, window=[0, 1])
            self.phase_polys.append(Polys(y_polys, dy_polys, u_polys))

    def refine_mesh(self):
        from .mesh_refinement import PattersonRaoMeshRefinement
        self.mesh_refinement = PattersonRaoMeshRefinement(self)
        return self.mesh_refinement

"""Mesh-refinement error on the GPU.

Mirror of ``pycollo/mesh_refinement.py:60-240`` up to and including the error
evaluation: the p+1 ("ph") mesh (``:75-86``), the solution interpolated to it
(``construct_x_ph`` ``:160-204``) and ``phase_mesh_error`` (``:206-240``).  The
reference rebuilds a CasADi graph of ``dy_ph`` for every mesh iteration and then
loops over sections, states and nodes in Python; here one engine is created on
the ph mesh with unit scaling (``V=1, r=0`` as ``:149-150``) and
``pcx_mesh_error`` (``include/pcx.h``) evaluates ``dy_ph`` at every ph node,
contracts it with the ph integration blocks and reduces the per-section maxima
on the device.  The refinement *decision* (``next_iteration_mesh``, ``:242+``)
is control plane and stays out of scope.
"""
from __future__ import annotations

import numpy as np

from . import engine as _engine
from .mesh import Mesh, PhaseMesh

DEFAULT_MESH_TOLERANCE = 1e-7
DEFAULT_MAX_MESH_ITERATIONS = 10
PATTERSON_RAO = "patterson-rao"


def create_ph_mesh(mesh, collocation_points_min=2, collocation_points_max=10):
    """``mesh_refinement.py:75-86``: same sections, ``N_k + 1`` nodes each."""
    phase_meshes = []
    for K, h_K, N_K in zip(mesh.K, mesh.h_K, mesh.N_K):
        phase_meshes.append(PhaseMesh(number_mesh_sections=K, mesh_section_sizes=h_K,
                                      number_mesh_section_nodes=np.asarray(N_K) + 1))
    return Mesh(mesh.quadrature, phase_meshes, collocation_points_min,
                collocation_points_max + 1)


class MeshErrorEvaluator:
    """One engine on the ph mesh of an iteration's mesh; reusable for any number
    of ``x_ph`` vectors (and ``batch`` of them per call)."""

    def __init__(self, ocp, mesh, batch=1, device=0, collocation_points_max=21,
                 shard=None, **structure_kwargs):
        """``shard=(rank, world_size)``: the error pass of one mesh split over the
        ranks by contiguous section ranges -- it is
        section-local, so there is no exchange; each rank fills the entries of its
        sections (``local_sections``) and leaves the others untouched."""
        from .backend import lower_problem
        from .parallel import shard_range
        self.ph_mesh = create_ph_mesh(mesh, 2, collocation_points_max)
        self.low = lower_problem(ocp, self.ph_mesh.p, **structure_kwargs)
        S = self.low.S
        self.engine = _engine.Engine(S, self.low.layouts, self.low.header,
                                     batch=batch, device=device)
        # user basis: V = 1, r = 0, W = 1, w = 1 (mesh_refinement.py:149-150)
        self.engine.set_scaling(np.ones(S.n_var_ocp), np.zeros(S.n_var_ocp),
                                np.ones(S.n_con_ocp), 1.0)
        self.local_sections = [(0, int(t.K)) for t in S.ph]
        if shard is not None:
            rank, world = shard
            lo, hi = shard_range(S.num_tiles, world, rank)
            self.engine.set_shard(lo, hi)
            secs = []
            for ip in range(len(S.ph)):
                sel = np.flatnonzero(S.tile_phase[lo:hi] == ip) + lo
                secs.append((int(S.tile_k0[sel].min()), int(S.tile_k1[sel].max()))
                            if len(sel) else (0, 0))
            self.local_sections = secs

    def __call__(self, x_ph):
        return self.engine.mesh_error_host(x_ph)

    def global_maximum(self, max_rel_per_phase):
        """Largest relative error over the whole mesh: the local maximum of this
        rank's sections, then one ``all_reduce(MAX)`` when a process group exists."""
        import torch
        import torch.distributed as dist
        local = max([float(np.max(m[lo:hi])) for m, (lo, hi) in
                     zip(max_rel_per_phase, self.local_sections) if hi > lo] + [0.0])
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dev = f"cuda:{self.engine.device}" if dist.get_backend() == "nccl" else "cpu"
            t = torch.tensor([local], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            local = float(t[0])
        return local


class PattersonRaoMeshRefinement:
    """``mesh_error()`` of ``pycollo/mesh_refinement.py:60-73`` for a ``Solution``."""

    def __init__(self, solution):
        self.sol = solution
        self.it = solution.it
        self.ocp = solution.ocp
        self.backend = solution.backend
        self.mesh_error()

    def mesh_error(self):
        s = self.ocp.settings
        self.evaluator = MeshErrorEvaluator(self.ocp, self.it.mesh, device=self.it.device,
                                            collocation_points_max=s.collocation_points_max)
        self.ph_mesh = self.evaluator.ph_mesh
        self.x_ph = self.construct_x_ph_device()
        res = self.evaluator(self.x_ph)
        self.absolute_mesh_errors = [r[0] for r in res]
        self.relative_mesh_errors = [r[1] for r in res]
        self.maximum_relative_mesh_errors = [r[2] for r in res]

    def construct_x_ph_device(self):
        """``mesh_refinement.py:160-204`` + ``solution_abc.py:60-142`` on the device
        (``pcx_refit_to_ph``): the per-section fits are applied as per-order
        matrices to the solution and its state derivatives."""
        it, sol = self.it, self.sol
        x_user = it.scaling.unscale_x(sol.x)
        dy = np.concatenate([np.ravel(p.dy) for p in sol.phase_data]) \
            if sol.phase_data else np.zeros(0)
        x_ph = it.create_engine().refit_to_ph_host(x_user, dy)
        self.y_ph, self.u_ph = [], []
        for t, irp in zip(self.evaluator.low.S.ph, self.backend.ir.phases):
            ny, nu = irp.n_y, irp.n_u
            blk = x_ph[t.x_off:t.x_off + (ny + nu) * t.N].reshape(ny + nu, t.N)
            self.y_ph.append(blk[:ny])
            self.u_ph.append(blk[ny:])
        return x_ph

    def construct_x_ph(self):
        """``mesh_refinement.py:160-204`` with the reference's numpy polynomial
        objects (host mirror; the product path is ``construct_x_ph_device``)."""
        parts, y_all, u_all = [], [], []
        for ip, p_data in enumerate(self.sol.phase_data):
            bnd = self.it.mesh.mesh_index_boundaries[ip]
            bnd_ph = self.ph_mesh.mesh_index_boundaries[ip]
            tau_ph = self.ph_mesh.tau[ip]
            polys = self.sol.phase_polys[ip]
            y_ph = self._to_ph(p_data.y, polys.y, bnd, bnd_ph, tau_ph)
            u_ph = self._to_ph(p_data.u, polys.u, bnd, bnd_ph, tau_ph)
            y_all.append(y_ph)
            u_all.append(u_ph)
            parts += [y_ph.ravel(), u_ph.ravel(), np.ravel(p_data.q), np.ravel(p_data.t)]
        parts.append(np.ravel(self.sol._s))
        return np.concatenate(parts), y_all, u_all

    @staticmethod
    def _to_ph(vals, polys, bnd, bnd_ph, tau_ph):
        out = np.zeros((vals.shape[0], len(tau_ph)))
        if vals.shape[0] == 0:
            return out
        out[:, bnd_ph] = vals[:, bnd]
        for i_var in range(vals.shape[0]):
            for i_k, (a, b) in enumerate(zip(bnd_ph[:-1], bnd_ph[1:])):
                out[i_var, a + 1:b] = polys[i_var, i_k](tau_ph[a + 1:b])
        return out

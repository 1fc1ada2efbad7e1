"""Temporal mesh of one phase and its collocation operators.

Host-side producer of the arrays the CUDA engine keeps device-resident
(SURVEY.md §8 row a10): the node abscissae ``tau`` on [-1, 1], the per-section
integration blocks ``I_k = A(N_k) * h_k``, the quadrature row ``W`` and the
section table ``(first node, N_k, h_k)``.

Layout follows ``pycollo/mesh.py:236-356``: a phase has ``K`` sections, section
``k`` has ``N_k`` nodes and shares its last node with section ``k+1``, so
``N = sum(N_k - 1) + 1``.  Defect rows are numbered section-major; row
``b_k + l`` (``l = 0..N_k-2``) belongs to node ``b_k + l + 1`` of section ``k``.

Naming: the reference stores the *integration* CSR as ``sI_matrix`` and the
*difference* CSR (+1 at the section start, -1 at the row's own node) as
``sA_matrix`` (``pycollo/mesh.py:332-335, 352-353``).  The same attribute names
are offered here so the reference's call sites read the same.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sparse

from .quadrature import Quadrature

TAU_0 = -1.0
TAU_F = 1.0


class PhaseMesh:
    """User-facing description of one phase's mesh (``pycollo/mesh.py:10-121``)."""

    def __init__(self, number_mesh_sections=10, mesh_section_sizes=None,
                 number_mesh_section_nodes=4):
        self.number_mesh_sections = int(number_mesh_sections)
        K = self.number_mesh_sections
        if mesh_section_sizes is None:
            sizes = np.ones(K) / K
        else:
            sizes = np.array(mesh_section_sizes, dtype=np.float64)
            if sizes.shape != (K,):
                raise ValueError(
                    f"Mesh section sizes must be an iterable of length {K} "
                    f"(i.e. matching the number of mesh sections).")
        self.mesh_section_sizes = sizes / sizes.sum()
        nodes = np.asarray(number_mesh_section_nodes)
        if nodes.ndim == 0:
            nodes = np.full(K, int(nodes), dtype=np.int64)
        else:
            nodes = nodes.astype(np.int64)
        if nodes.shape != (K,):
            raise ValueError(
                f"Number of mesh section nodes must be an interable of length "
                f"{K} (i.e. matching the number of mesh sections).")
        self.number_mesh_section_nodes = nodes


class PhaseMeshData:
    """Numeric mesh of one phase (what ``Mesh.generate_single_phase`` returns)."""

    def __init__(self, quadrature: Quadrature, phase_mesh: PhaseMesh,
                 collocation_points_min=2, collocation_points_max=10):
        nodes = np.asarray(phase_mesh.number_mesh_section_nodes)
        low = np.flatnonzero(nodes < collocation_points_min)
        high = np.flatnonzero(nodes > collocation_points_max)
        first = min([int(a[0]) for a in (low, high) if a.size], default=None)
        if first is not None:
            n, k = nodes[first], first
            if n < collocation_points_min:
                raise ValueError(
                    f"The number of collocation points, {n}, in mesh section "
                    f"{k} must be greater than or equal to "
                    f"{collocation_points_min}.")
            raise ValueError(
                f"The number of collocation points, {n}, in mesh section "
                f"{k} must be less than or equal to "
                f"{collocation_points_max}.")
        self.quadrature = quadrature
        self.K = phase_mesh.number_mesh_sections
        self.N_K = nodes.copy()

        # section boundaries by running sum, as pycollo/mesh.py:248-253
        bounds = [TAU_0]
        for frac in phase_mesh.mesh_section_sizes:
            bounds.append(bounds[-1] + (TAU_F - TAU_0) * frac)
        bounds = np.array(bounds)

        # Sections of one order are mapped together (10^6-node meshes: a Python loop over
        # the sections took seconds per phase); element by element the arithmetic is the
        # per-section statement of pycollo/mesh.py:283-326 -- stretch * points + shift,
        # weights * h_k -- so the arrays equal the loop's (and the reference's) bit for bit.
        self.mesh_index_boundaries = np.concatenate(
            [[0], np.cumsum(nodes - 1)]).astype(np.int64)
        start = self.mesh_index_boundaries[:-1]
        self.N = int(self.mesh_index_boundaries[-1]) + 1
        self.tau = np.empty(self.N)
        self.tau[-1] = TAU_F
        groups = [(int(n), np.flatnonzero(nodes == n)) for n in np.unique(nodes)]
        for n, idx in groups:
            pts = quadrature.quadrature_point(n)
            stretch = 0.5 * (bounds[idx + 1] - bounds[idx])
            shift = 0.5 * (bounds[idx] + bounds[idx + 1])
            mapped = stretch[:, None] * pts[None, :] + shift[:, None]
            self.tau[start[idx][:, None] + np.arange(n - 1)[None, :]] = mapped[:, :-1]
        self.h = np.diff(self.tau)
        self.h_K = np.diff(self.tau[self.mesh_index_boundaries])
        self.num_c_defect_per_y = int(self.mesh_index_boundaries[-1])

        # quadrature row (a shared end node receives two contributions: their sum does not
        # depend on the order); the dense per-section integration blocks are formed on demand
        self.W_matrix = np.zeros(self.N)
        for n, idx in groups:
            contrib = quadrature.quadrature_weight(n)[None, :] * self.h_K[idx][:, None]
            np.add.at(self.W_matrix, start[idx][:, None] + np.arange(n)[None, :], contrib)
        self._I_blocks = None

    @property
    def I_blocks(self):
        """Per-section integration blocks ``ButcherRows(N_k) * h_k`` (``pycollo/mesh.py:300``)."""
        if self._I_blocks is None:
            self._I_blocks = [self.quadrature.A_matrix(int(n)) * hk
                              for n, hk in zip(self.N_K, self.h_K)]
        return self._I_blocks

    # -- CSR views with the reference's attribute names -----------------
    @property
    def sI_matrix(self):
        """Integration operator, (N-1) x N CSR, explicit zeros kept."""
        rows, cols, vals = [], [], []
        for k, blk in enumerate(self.I_blocks):
            b = int(self.mesh_index_boundaries[k])
            n = int(self.N_K[k])
            r, c = np.meshgrid(np.arange(b, b + n - 1), np.arange(b, b + n),
                               indexing="ij")
            rows.append(r.ravel())
            cols.append(c.ravel())
            vals.append(blk.ravel())
        return sparse.coo_matrix(
            (np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
            shape=(self.N - 1, self.N)).tocsr()

    @property
    def sA_matrix(self):
        """Difference operator: +1 at the section start, -1 at the row's node."""
        rows, cols, vals = [], [], []
        for k in range(self.K):
            b = int(self.mesh_index_boundaries[k])
            n = int(self.N_K[k])
            r = np.arange(b, b + n - 1)
            rows += [r, r]
            cols += [np.full(n - 1, b), r + 1]
            vals += [np.ones(n - 1), -np.ones(n - 1)]
        return sparse.coo_matrix(
            (np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
            shape=(self.N - 1, self.N)).tocsr()

    @classmethod
    def from_arrays(cls, quadrature, N_K, h_K, I_blocks, W_matrix, tau=None):
        """Adopt mesh operators produced elsewhere (e.g. by the reference's own
        ``Mesh`` object) without regenerating them -- the drop-in path."""
        self = cls.__new__(cls)
        self.quadrature = quadrature
        self.N_K = np.asarray(N_K, dtype=np.int64)
        self.K = int(self.N_K.size)
        self.mesh_index_boundaries = np.concatenate(
            [[0], np.cumsum(self.N_K - 1)]).astype(np.int64)
        self.N = int(self.mesh_index_boundaries[-1]) + 1
        self.num_c_defect_per_y = self.N - 1
        self.h_K = np.asarray(h_K, dtype=np.float64)
        self._I_blocks = [np.asarray(b, dtype=np.float64) for b in I_blocks]
        self.W_matrix = np.asarray(W_matrix, dtype=np.float64)
        self.tau = None if tau is None else np.asarray(tau, dtype=np.float64)
        self.h = None if tau is None else np.diff(self.tau)
        return self

    @classmethod
    def from_reference_csr(cls, quadrature, N_K, tau, sI_csr, W_matrix):
        """Build from the reference's ``mesh.sI_matrix[p]`` CSR and ``W_matrix``."""
        N_K = np.asarray(N_K, dtype=np.int64)
        bnd = np.concatenate([[0], np.cumsum(N_K - 1)]).astype(np.int64)
        dense_rows = sI_csr.tocsr()
        blocks = []
        for k, n in enumerate(N_K):
            b = int(bnd[k])
            blocks.append(np.asarray(
                dense_rows[b:b + n - 1, b:b + n].todense(), dtype=np.float64))
        tau = np.asarray(tau, dtype=np.float64)
        return cls.from_arrays(quadrature, N_K, np.diff(tau[bnd]), blocks,
                               W_matrix, tau)


class Mesh:
    """All phases' meshes; list-valued attributes as ``pycollo/mesh.py:204-234``."""

    def __init__(self, quadrature, phase_meshes, collocation_points_min=2,
                 collocation_points_max=10):
        self.quadrature = quadrature
        self.p = [pm if isinstance(pm, PhaseMeshData) else
                  PhaseMeshData(quadrature, pm, collocation_points_min,
                                collocation_points_max)
                  for pm in phase_meshes]

    def _collect(self, name):
        return [getattr(p, name) for p in self.p]

    tau = property(lambda self: self._collect("tau"))
    h = property(lambda self: self._collect("h"))
    N = property(lambda self: self._collect("N"))
    K = property(lambda self: self._collect("K"))
    N_K = property(lambda self: self._collect("N_K"))
    h_K = property(lambda self: self._collect("h_K"))
    mesh_index_boundaries = property(
        lambda self: self._collect("mesh_index_boundaries"))
    num_c_defect_per_y = property(
        lambda self: self._collect("num_c_defect_per_y"))
    W_matrix = property(lambda self: self._collect("W_matrix"))
    sI_matrix = property(lambda self: self._collect("sI_matrix"))
    sA_matrix = property(lambda self: self._collect("sA_matrix"))

"""Sparsity structure, scatter recipes and tiling of the NLP callbacks.

Built once per mesh on the host (integer work only), uploaded once, never
re-read from HBM at full size during an evaluation:

* the x / c layouts of ``pycollo/backend.py:1433-1457, 1551-1563`` and
  ``pycollo/iteration.py:196-342``;
* the Jacobian pattern in CasADi CCS order -- sorted by (column, row) -- as
  returned by ``Casadi.evaluate_G_structure`` (``backend.py:1747-1761``) and the
  upper-triangular CCS pattern of the Lagrangian Hessian (``backend.py:1693``
  ``nlpsol`` convention), both constructed *in order* (no sort) so 10^6-node
  meshes set up in seconds;
* per section-type "recipes": for every value slot of the columns a mesh
  section owns, which staged node derivative, which quadrature coefficient and
  which constant produce it.  A section type is (N_prev, N_k, last?) so a
  uniform mesh has three types and the recipe table stays L1/L2-resident;
* tiles: contiguous section ranges, one CTA each, sized so the grid is a
  multiple of the SM count;
* the "border": the O(1) slots that are not node-local (integral rows,
  endpoint rows, q/t/s corner of H, the two end nodes of H) written by the
  last-arriving CTA as a tiny sparse linear map from a border-value vector.

Exact-zero quadrature coefficients are pruned from the pattern by default
(Radau's last integration column and last weight, ``quadrature.py:133,138``),
the single switch ``prune`` (SURVEY.md §7 "hard parts").
"""
from __future__ import annotations

from dataclasses import dataclass, field

import os

import numpy as np


def sym_free(expr):
    """Free symbols of a sympy expression (empty for plain numbers)."""
    return getattr(expr, "free_symbols", ())


# recipe word bit layout (must match csrc/pcx_params.h).  Low word: staged row
# (9 bits), quadrature-table index (13), constant index (7), three flags; high
# word: variable (8), slot inside the variable's period (18), node inside the
# section (5 bits: sections of up to 20 nodes, Settings.collocation_points_max).
RC_E_BITS, RC_B_BITS, RC_M_BITS, RC_C_BITS = 9, 13, 5, 7
RC_E_SHIFT = 0
RC_B_SHIFT = RC_E_SHIFT + RC_E_BITS
RC_C_SHIFT = RC_B_SHIFT + RC_B_BITS           # 22
RC_PREV_BIT = RC_C_SHIFT + RC_C_BITS          # 29
RC_PLAIN_BIT = RC_PREV_BIT + 1                # 30
RC_SKIP_BIT = RC_PLAIN_BIT + 1                # 31
RC_VAR_SHIFT = 32                             # variable index (8 bits)
RC_LOCAL_SHIFT = 40                           # slot inside the variable's period
RC_LOCAL_BITS = 18
RC_M_SHIFT = RC_LOCAL_SHIFT + RC_LOCAL_BITS   # 58
MAX_SECTION_NODES = 20

# CTAs the kernels are compiled to keep resident per SM on large meshes
RESIDENT_CTAS = 6
LARGE_MESH_THREADS = 128
SMALL_BODY_THREADS = 192


def resident_ctas(threads):
    """CTAs per SM the tiling and the register cap aim at: 24 warps per SM."""
    return {160: 5, 192: 4, 256: 3}.get(int(threads), RESIDENT_CTAS if threads <= 128 else 3)

# border-map groups
GRP_C, GRP_G, GRP_H, GRP_J, GRP_GRAD = 0, 1, 2, 3, 4

# scale-vector references: products of entries of [1, V_ocp.., W_ocp.., w]
@dataclass
class ScaleIndex:
    n_var: int
    n_con: int

    def one(self):
        return 0

    def V(self, i):
        return 1 + i

    def W(self, i):
        return 1 + self.n_var + i

    def w(self):
        return 1 + self.n_var + self.n_con

    def vector(self, V_ocp, W_ocp, w):
        return np.concatenate([[1.0], V_ocp, W_ocp, [w]]).astype(np.float64)


class Products:
    """List of ``const * prod(scales[refs])`` terms evaluated on set_scaling."""

    def __init__(self):
        self.const = []
        self.refs = []

    def add(self, const, *refs):
        refs = list(refs)[:3] + [0] * (3 - len(refs))
        self.const.append(float(const))
        self.refs.append(refs)
        return len(self.const) - 1

    def evaluate(self, scales):
        if not self.const:
            return np.zeros(0)
        refs = np.asarray(self.refs, dtype=np.int64)
        return np.asarray(self.const) * scales[refs[:, 0]] * scales[refs[:, 1]] \
            * scales[refs[:, 2]]

    def __len__(self):
        return len(self.const)


@dataclass
class PhaseTables:
    """Everything integer/structural about one phase."""
    N: int = 0
    K: int = 0
    x_off: int = 0
    c_off: int = 0
    dy_off: int = 0
    W_off: int = 0               # offset of this phase in W_ocp
    V_off: dict = field(default_factory=dict)   # OCP-level offsets y,u,q,t
    q_col: int = 0
    t_cols: list = field(default_factory=list)  # x index of t0/tF or -1
    sec_node: np.ndarray = None
    sec_order: np.ndarray = None
    sec_h: np.ndarray = None
    sec_type: np.ndarray = None
    sec_off: int = 0             # offset into the global section arrays
    gsec_ptr: np.ndarray = None  # (NV, K+1) int64 G slot of each section period
    type_ids: list = field(default_factory=list)
    # row-chunk bases in G
    g_tcol_base: dict = field(default_factory=dict)   # (tk, i) -> slot
    g_scol_base: dict = field(default_factory=dict)   # (j, e) -> slot
    # H bases
    h_reg_base: np.ndarray = None     # (NV,) slot of column (b, 1)
    h_nA: np.ndarray = None           # (NV,) entries per regular column
    h_pair_pos: np.ndarray = None     # (NH2VV,) position inside its column
    h_t_base: dict = field(default_factory=dict)      # (tk, a) -> slot of row (a,1)
    h_s_base: dict = field(default_factory=dict)      # (a, j) -> slot of row (a,1)
    bv_irr: list = field(default_factory=list)        # BV offset for node 0 / N-1
    red_off: int = 0


class NLPStructure:
    def __init__(self, ir, phase_derivs, point_derivs, meshes, *, prune=True,
                 sm_count=148, threads=None, max_tile_nodes=None,
                 smem_budget=96 * 1024, tiles_per_sm=None):
        self.ir = ir
        self.pd = phase_derivs
        self.pt = point_derivs
        self.meshes = meshes
        self.prune = bool(prune)
        if threads is None:
            # CTA size: one thread per node of a tile; small problems (multi-start
            # sweeps of ~31-node meshes) get a CTA no wider than their mesh
            nmax = max(int(m.N) for m in meshes)
            # large meshes: 128 threads x 6 CTAs/SM; small expression bodies (the
            # ones the kernel parks in registers, <= 40 results per node: cart-pole)
            # run 4 % faster with 192 x 4 -- fewer, fatter tiles amortise the
            # per-tile prologue -- while larger bodies (robot, Delta III) lose
            # 4-11 % there (measured, tools/knobs.py / tools/threads_ab.sh)
            nout = max(pd.NF + len(pd.d1v) + len(pd.d1s) + len(pd.h2vv) + len(pd.h2vs)
                       + len(pd.h2ss) + len(pd.htv) + len(pd.hts) for pd in phase_derivs)
            large = SMALL_BODY_THREADS if nout <= 40 else LARGE_MESH_THREADS
            threads = 32 if nmax <= 32 else (64 if nmax <= 64 else
                                             int(os.environ.get("PCX_THREADS", large)))
        self.threads = int(threads)
        self.P = len(ir.phases)
        self.NS = ir.n_s
        self.NB = ir.n_b
        self._layout()
        self._quadrature_tables()
        self._scales_setup()
        self._build_G()
        self._build_H()
        if os.environ.get("PCX_SMEM_BUDGET"):            # experiment knob: bytes staged per tile
            smem_budget = int(os.environ["PCX_SMEM_BUDGET"])
        # experiment knob (tools/knobs.py): tiles per SM on large meshes
        if tiles_per_sm is None and os.environ.get("PCX_TILES_PER_SM"):
            tiles_per_sm = int(os.environ["PCX_TILES_PER_SM"])
        self._build_tiles(sm_count, max_tile_nodes, smem_budget, tiles_per_sm)
        self._build_border()

    # ------------------------------------------------------------ layout --
    def _layout(self):
        self.ph = []
        xo = co = dyo = wo = vo = so = 0
        for ph, pd, mesh in zip(self.ir.phases, self.pd, self.meshes):
            N = int(mesh.N)
            t = PhaseTables(N=N, K=int(mesh.K), x_off=xo, c_off=co, dy_off=dyo,
                            W_off=wo, sec_off=so)
            t.V_off = dict(y=vo, u=vo + ph.n_y, q=vo + ph.n_y + ph.n_u,
                           t=vo + ph.n_y + ph.n_u + ph.n_q)
            t.q_col = xo + pd.NV * N
            tc = t.q_col + ph.n_q
            t.t_cols = []
            t.t_ocp = []
            k = 0
            for needed in ph.t_needed:
                if needed:
                    t.t_cols.append(tc + k)
                    t.t_ocp.append(t.V_off["t"] + k)
                    k += 1
                else:
                    t.t_cols.append(-1)
                    t.t_ocp.append(-1)
            t.t_const = [float(ph.t0) if not ph.t_needed[0] else 0.0,
                         float(ph.tF) if not ph.t_needed[1] else 0.0]
            xo += pd.NV * N + ph.n_q + ph.n_t
            co += ph.n_y * (N - 1) + ph.n_p * N + ph.n_q
            dyo += ph.n_y * N
            wo += ph.n_y + ph.n_p + ph.n_q
            vo += ph.n_y + ph.n_u + ph.n_q + ph.n_t
            so += int(mesh.K)
            self.ph.append(t)
        self.s_off = xo
        self.num_x = xo + self.NS
        self.b_off = co
        self.num_c = co + self.NB
        self.num_dy = dyo
        self.Wb_off = wo
        self.n_con_ocp = wo + self.NB
        self.s_ocp_off = vo
        self.n_var_ocp = vo + self.NS
        self.total_sections = so
        # point variables: x index and OCP-level scale index, in pts order
        pidx, pV = [], []
        for ph, t in zip(self.ir.phases, self.ph):
            for i in range(ph.n_y):
                pidx += [t.x_off + i * t.N, t.x_off + i * t.N + t.N - 1]
                pV += [t.V_off["y"] + i] * 2
            for i in range(ph.n_q):
                pidx.append(t.q_col + i)
                pV.append(t.V_off["q"] + i)
            for k in range(ph.n_t):
                pidx.append(t.q_col + ph.n_q + k)
                pV.append(t.V_off["t"] + k)
        for j in range(self.NS):
            pidx.append(self.s_off + j)
            pV.append(self.s_ocp_off + j)
        self.pt_x = np.asarray(pidx, dtype=np.int64)
        self.pt_V = np.asarray(pV, dtype=np.int64)
        assert len(self.pt_x) == len(self.pt.pts)
        assert np.all(np.diff(self.pt_x) > 0)

    def _quadrature_tables(self):
        orders = sorted({int(n) for m in self.meshes for n in m.N_K})
        quad = self.meshes[0].quadrature
        btab = [0.0, 1.0]            # [0] = 0.0, [1] = 1.0 sentinels
        self.a_off, self.w_off = {}, {}
        self.a_nz, self.w_nz = {}, {}
        for n in orders:
            A = np.asarray(quad.A_matrix(n), dtype=np.float64)
            wv = np.asarray(quad.quadrature_weight(n), dtype=np.float64)
            self.a_off[n] = len(btab)
            btab.extend(A.ravel().tolist())
            self.w_off[n] = len(btab)
            btab.extend(wv.tolist())
            self.a_nz[n] = (A != 0.0) if self.prune else np.ones_like(A, bool)
            self.w_nz[n] = (wv != 0.0) if self.prune else np.ones_like(wv, bool)
        self.btab = np.asarray(btab, dtype=np.float64)
        self.orders = orders
        if len(self.btab) >= (1 << RC_B_BITS):
            raise ValueError("too many distinct section orders for the recipe "
                             f"word (quadrature table exceeds {1 << RC_B_BITS} entries)")
        omax = max(orders)
        self.order_a_off = np.zeros(omax + 1, dtype=np.int32)
        self.order_w_off = np.zeros(omax + 1, dtype=np.int32)
        for n in orders:
            self.order_a_off[n] = self.a_off[n]
            self.order_w_off[n] = self.w_off[n]

    def _scales_setup(self):
        self.sidx = ScaleIndex(self.n_var_ocp, self.n_con_ocp)

    # ----------------------------------------------------------------- G --
    def _section_types(self, ip):
        """Assign a type to every section of phase ip and build its recipes."""
        ph, pd, mesh, t = self.ir.phases[ip], self.pd[ip], self.meshes[ip], self.ph[ip]
        NK = np.asarray(mesh.N_K, dtype=np.int64)
        K = len(NK)
        prev = np.concatenate([[0], NK[:-1]])
        last = np.zeros(K, dtype=np.int64)
        last[-1] = 1
        key = prev * 64 + NK * 2 + last
        uniq, inv = np.unique(key, return_inverse=True)
        t.sec_type_local = inv.astype(np.int64)
        t.type_keys = [(int(k // 64), int((k % 64) // 2), int(k % 2)) for k in uniq]

    def _wq_nonzero(self, n_prev, n_k, mloc, is_last_node):
        if not self.prune:
            return True
        nz = bool(self.w_nz[n_k][mloc])
        if mloc == 0 and n_prev > 0:
            nz = nz or bool(self.w_nz[n_prev][n_prev - 1])
        return nz

    def _type_recipes(self, ip, n_prev, n_k, is_last):
        """Per var: list of entries for the columns a section of this type owns.

        entry = (mloc, rowkind, idx, l, prev, stage_row, bidx, cidx, plain, skip)
        rowkind: 0 defect(i=idx, row l of prev/current section), 1 path(j),
                 2 integral(i), 3 endpoint(k)
        """
        ph, pd = self.ir.phases[ip], self.pd[ip]
        NY, NP, NV = pd.NY, pd.NP, pd.NV
        d1v_index = {ea: k for k, ea in enumerate(pd.d1v)}
        ptd = self.pt
        # endpoint dependencies: point index of y_a(t0) / y_a(tF) in pts order
        pt_base = 0
        for q in range(ip):
            pq = self.ir.phases[q]
            pt_base += 2 * pq.n_y + pq.n_q + pq.n_t
        b_dep = {}
        for (e, a) in ptd.d1:
            if e >= 1:
                b_dep.setdefault(a, []).append(e - 1)
        owned = list(range(n_k - 1)) + ([n_k - 1] if is_last else [])
        out = []
        for a in range(NV):
            ents = []
            for mloc in owned:
                for i in range(NY):
                    st = d1v_index.get((i, a))
                    hasI = st is not None
                    hasD = (a == i)
                    if mloc == 0 and n_prev > 0:
                        for l in range(n_prev - 1):
                            Inz = hasI and bool(self.a_nz[n_prev][l, n_prev - 1])
                            Dnz = hasD and l == n_prev - 2
                            if Inz or Dnz:
                                ents.append((mloc, 0, i, l, 1,
                                             st + 1 if Inz else 0,
                                             self.a_off[n_prev] + l * n_prev
                                             + n_prev - 1,
                                             2 + 2 * i if Dnz else 0, 0, 0))
                    for l in range(n_k - 1):
                        Inz = hasI and bool(self.a_nz[n_k][l, mloc])
                        Dnz = hasD and (mloc == 0 or mloc == l + 1)
                        if Inz or Dnz:
                            c = 0
                            if Dnz:
                                c = 1 + 2 * i if mloc == 0 else 2 + 2 * i
                            ents.append((mloc, 0, i, l, 0, st + 1 if Inz else 0,
                                         self.a_off[n_k] + l * n_k + mloc, c,
                                         0, 0))
                for j in range(NP):
                    st = d1v_index.get((NY + j, a))
                    if st is not None:
                        ents.append((mloc, 1, j, 0, 0, st + 1, 1, 0, 1, 0))
                wq_on = self._wq_nonzero(n_prev, n_k, mloc, False)
                for i in range(pd.NQ):
                    st = d1v_index.get((NY + NP + i, a))
                    if st is not None and wq_on:
                        ents.append((mloc, 2, i, 0, 0, st + 1, 1, 0, 1, 0))
                if a < NY:
                    pt = None
                    if mloc == 0 and n_prev == 0:
                        pt = pt_base + 2 * a
                    elif is_last and mloc == n_k - 1:
                        pt = pt_base + 2 * a + 1
                    if pt is not None:
                        for k in b_dep.get(pt, []):
                            ents.append((mloc, 3, k, 0, 0, 0, 0, 0, 1, 1))
            out.append(ents)
        return out

    def _build_G(self):
        sidx = self.sidx
        self.types = []            # global list of (ip, n_prev, n_k, is_last)
        self.type_recipes = []     # per type: list per var of entries
        self.recipe_words = []
        self.type_var_off = []     # (ntypes, NVMAX+1) offsets into recipe_words
        self.NVMAX = max(pd.NV for pd in self.pd)
        slot = 0
        self.g_border = []         # (slot, row, col, descriptor)
        g_rows_parts, g_cols_parts, g_slot_parts = [], [], []
        g_thunks = []              # deferred pattern blocks: () -> (rows, cols, slots)
        words = []
        for ip, (ph, pd, mesh, t) in enumerate(zip(self.ir.phases, self.pd,
                                                   self.meshes, self.ph)):
            self._section_types(ip)
            NY, NP, NQ, NV, N, K = pd.NY, pd.NP, pd.NQ, pd.NV, t.N, t.K
            if len(pd.d1v) + 1 >= (1 << RC_E_BITS):
                raise ValueError("too many first-derivative entries per node "
                                 "for the recipe word")
            if 2 + 2 * NY >= (1 << RC_C_BITS):
                raise ValueError("too many states for the recipe word")
            tid0 = len(self.types)
            t.type_ids = []
            for (n_prev, n_k, is_last) in t.type_keys:
                rec = self._type_recipes(ip, n_prev, n_k, is_last)
                self.types.append((ip, n_prev, n_k, is_last))
                self.type_recipes.append(rec)
                offs = []
                for a in range(self.NVMAX):
                    offs.append(len(words))
                    if a < NV:
                        for (mloc, rk, idx, l, prev, st, bidx, cidx, plain,
                             skip) in rec[a]:
                            local = len(words) - offs[-1]
                            if local >= (1 << RC_LOCAL_BITS) or mloc >= (1 << RC_M_BITS):
                                raise ValueError("section too large for the recipe word")
                            words.append(
                                (st << RC_E_SHIFT) | (bidx << RC_B_SHIFT)
                                | (mloc << RC_M_SHIFT) | (cidx << RC_C_SHIFT)
                                | (prev << RC_PREV_BIT) | (plain << RC_PLAIN_BIT)
                                | (skip << RC_SKIP_BIT)
                                | (a << RC_VAR_SHIFT) | (local << RC_LOCAL_SHIFT))
                offs.append(len(words))
                self.type_var_off.append(offs)
                t.type_ids.append(len(self.types) - 1)
            t.sec_type = (t.sec_type_local + tid0).astype(np.int32)
            t.sec_node = np.asarray(mesh.mesh_index_boundaries, dtype=np.int64)
            t.sec_order = np.asarray(mesh.N_K, dtype=np.int32)
            t.sec_h = np.asarray(mesh.h_K, dtype=np.float64)
            plen = np.array([[len(self.type_recipes[tid][a]) for a in range(NV)]
                             for tid in t.type_ids], dtype=np.int64)  # (ntype, NV)
            t.gsec_ptr = np.zeros((NV, K + 1), dtype=np.int64)
            for a in range(NV):
                lens = plen[t.sec_type_local, a]
                t.gsec_ptr[a, 0] = slot
                np.cumsum(lens, out=t.gsec_ptr[a, 1:])
                t.gsec_ptr[a, 1:] += slot
                slot = int(t.gsec_ptr[a, K])
                # expand rows/cols, vectorised per type
                for lt, tid in enumerate(t.type_ids):
                    ents = self.type_recipes[tid][a]
                    if not ents:
                        continue
                    ks = np.flatnonzero(t.sec_type_local == lt)
                    E = np.asarray(ents, dtype=np.int64)
                    # the (row, col, slot) triplets of these sections are formed when the
                    # pattern is asked for (G_structure): at 10^6 nodes they are 10^8
                    # entries that an evaluation-only engine never needs
                    g_thunks.append(self._recipe_pattern(t, a, ks, E, NY, NP, N))
                    # endpoint (skip) entries go to the border map
                    for ie, ent in enumerate(ents):
                        if ent[9]:
                            for k_sec in ks:
                                self.g_border.append(
                                    (int(t.gsec_ptr[a, k_sec]) + ie,
                                     ("b_d1", int(ent[2]),
                                      int(t.x_off + a * N + t.sec_node[k_sec] + ent[0]))))
            # ---- border columns of this phase: q, then free t ----
            for i in range(NQ):
                col = t.q_col + i
                r_int = t.c_off + NY * (N - 1) + NP * N + i
                self._g_explicit(g_rows_parts, g_cols_parts, g_slot_parts,
                                 slot, r_int, col)
                self.g_border.append((slot, ("const", sidx.W(t.W_off + NY + NP + i),
                                             sidx.V(t.V_off["q"] + i))))
                slot += 1
                slot = self._g_endpoint_rows(g_rows_parts, g_cols_parts,
                                             g_slot_parts, slot, col)
            for tk in range(2):
                col = t.t_cols[tk]
                if col < 0:
                    continue
                for i in range(NY):
                    if not pd.fn_nonzero[i]:
                        continue
                    t.g_tcol_base[(tk, i)] = slot
                    r0 = t.c_off + i * (N - 1)
                    g_rows_parts.append(r0 + np.arange(N - 1))
                    g_cols_parts.append(np.full(N - 1, col))
                    g_slot_parts.append(slot + np.arange(N - 1))
                    slot += N - 1
                for i in range(NQ):
                    if not pd.fn_nonzero[NY + NP + i]:
                        continue
                    r_int = t.c_off + NY * (N - 1) + NP * N + i
                    self._g_explicit(g_rows_parts, g_cols_parts, g_slot_parts,
                                     slot, r_int, col)
                    self.g_border.append((slot, ("int_t", ip, i, tk)))
                    slot += 1
                slot = self._g_endpoint_rows(g_rows_parts, g_cols_parts,
                                             g_slot_parts, slot, col)
        # ---- static parameter columns ----
        for j in range(self.NS):
            col = self.s_off + j
            for ip, (pd, t) in enumerate(zip(self.pd, self.ph)):
                NY, NP, NQ, N = pd.NY, pd.NP, pd.NQ, t.N
                deps = {e for e, jj in pd.d1s if jj == j}
                for i in range(NY):
                    if i in deps:
                        t.g_scol_base[(j, i)] = slot
                        g_rows_parts.append(t.c_off + i * (N - 1) + np.arange(N - 1))
                        g_cols_parts.append(np.full(N - 1, col))
                        g_slot_parts.append(slot + np.arange(N - 1))
                        slot += N - 1
                for jj in range(NP):
                    if NY + jj in deps:
                        t.g_scol_base[(j, NY + jj)] = slot
                        g_rows_parts.append(t.c_off + NY * (N - 1) + jj * N
                                            + np.arange(N))
                        g_cols_parts.append(np.full(N, col))
                        g_slot_parts.append(slot + np.arange(N))
                        slot += N
                for i in range(NQ):
                    if NY + NP + i in deps:
                        r_int = t.c_off + NY * (N - 1) + NP * N + i
                        self._g_explicit(g_rows_parts, g_cols_parts,
                                         g_slot_parts, slot, r_int, col)
                        self.g_border.append((slot, ("int_s", ip, i, j)))
                        slot += 1
            slot = self._g_endpoint_rows(g_rows_parts, g_cols_parts, g_slot_parts,
                                         slot, col)
        self.nnz_g = slot
        self.recipe_words = np.asarray(words, dtype=np.uint64)
        self.type_var_off = np.asarray(self.type_var_off, dtype=np.int32)
        self._g_parts = (g_rows_parts, g_cols_parts, g_slot_parts, g_thunks)
        self._G_rows = self._G_cols = None

    def G_constant_ranges(self, min_len=4096):
        """Maximal runs ``[lo, hi)`` of Jacobian value slots that do NOT depend on the
        iterate (only on the mesh and the scaling), at least ``min_len`` slots long.

        A slot of a (y, u) column is constant when it holds no staged derivative (the
        +-1 entries of the difference operator) or when its derivative expression is a
        number and nothing iterate-dependent multiplies it: with fixed phase times
        ``dq/dt = qd`` contributes ``h/2 * I[l, m]`` -- a constant of the mesh.  Whole
        variable blocks are constant for such states (cart-pole: the q1 and q1d
        columns, 20 % of the Jacobian), and a host that keeps its value array between
        evaluations need not fetch them again (``PCX_EVAL_CONST_RESIDENT``).
        Slots outside the (y, u) columns, entries written by the border pass and
        everything in a phase with a free time are treated as iterate-dependent."""
        mask = np.zeros(self.nnz_g, dtype=bool)
        for ip, (pd, t) in enumerate(zip(self.pd, self.ph)):
            free_time = t.t_cols[0] >= 0 or t.t_cols[1] >= 0
            number = [len(sym_free(e)) == 0 for e in pd.d1v_expr]
            for a in range(pd.NV):
                for lt, tid in enumerate(t.type_ids):
                    ents = self.type_recipes[tid][a]
                    if not ents:
                        continue
                    flags = np.zeros(len(ents), dtype=bool)
                    for ie, ent in enumerate(ents):
                        st, skip = ent[5], ent[9]
                        if skip:
                            continue
                        if st == 0:
                            flags[ie] = True
                        else:
                            fam = pd.fam[pd.d1v[st - 1][0]]
                            flags[ie] = number[st - 1] and (fam == "p" or not free_time)
                    if not flags.any():
                        continue
                    ks = np.flatnonzero(t.sec_type_local == lt)
                    sl = t.gsec_ptr[a, ks][:, None] + np.arange(len(ents))[None, :]
                    mask[sl[:, flags].ravel()] = True
        if not mask.any():
            return np.zeros((0, 2), dtype=np.int64)
        d = np.diff(np.concatenate([[0], mask.view(np.int8), [0]]))
        lo, hi = np.flatnonzero(d == 1), np.flatnonzero(d == -1)
        keep = (hi - lo) >= int(min_len)
        return np.stack([lo[keep], hi[keep]], axis=1).astype(np.int64)

    def _g_explicit(self, rp, cp, sp, slot, row, col):
        rp.append(np.array([row], dtype=np.int64))
        cp.append(np.array([col], dtype=np.int64))
        sp.append(np.array([slot], dtype=np.int64))

    def _g_endpoint_rows(self, rp, cp, sp, slot, col):
        """Endpoint-constraint rows depending on the point variable at x index col."""
        hit = np.flatnonzero(self.pt_x == col)
        if len(hit) == 0:
            return slot
        a = int(hit[0])
        for (e, aa) in self.pt.d1:
            if aa == a and e >= 1:
                self._g_explicit(rp, cp, sp, slot, self.b_off + e - 1, col)
                self.g_border.append((slot, ("b_d1", e - 1, col)))
                slot += 1
        return slot

    def _recipe_pattern(self, t, a, ks, E, NY, NP, N):
        """Deferred (rows, cols, slots) of variable ``a``'s recipe entries ``E`` in the
        sections ``ks`` of one type (vectorised per type)."""
        b_off = self.b_off

        def make():
            mloc, rk, idx, l, prev = E[:, 0], E[:, 1], E[:, 2], E[:, 3], E[:, 4]
            b_k = t.sec_node[ks][:, None]
            b_prev = t.sec_node[np.maximum(ks - 1, 0)][:, None]
            node = b_k + mloc[None, :]
            rows = np.where(
                rk[None, :] == 0,
                t.c_off + idx[None, :] * (N - 1)
                + np.where(prev[None, :] == 1, b_prev, b_k) + l[None, :],
                np.where(
                    rk[None, :] == 1,
                    t.c_off + NY * (N - 1) + idx[None, :] * N + node,
                    np.where(rk[None, :] == 2,
                             t.c_off + NY * (N - 1) + NP * N + idx[None, :],
                             b_off + idx[None, :])))
            cols = t.x_off + a * N + node
            sl = t.gsec_ptr[a, ks][:, None] + np.arange(E.shape[0])[None, :]
            return rows.ravel(), cols.ravel(), sl.ravel()
        return make

    def G_structure(self):
        if self._G_rows is None:
            rp, cp, sp, thunks = self._g_parts
            rows = np.empty(self.nnz_g, dtype=np.int64)
            cols = np.empty(self.nnz_g, dtype=np.int64)
            count = 0
            if rp:
                sl = np.concatenate(sp)
                rows[sl] = np.concatenate(rp)
                cols[sl] = np.concatenate(cp)
                count += len(sl)
            for make in thunks:
                r, c, sl = make()
                rows[sl] = r
                cols[sl] = c
                count += len(sl)
            assert count == self.nnz_g, (count, self.nnz_g)
            self._G_rows, self._G_cols = rows, cols
            self._g_parts = None
        return self._G_rows, self._G_cols

    # ----------------------------------------------------------------- H --
    def _build_H(self):
        """Upper-triangular CCS pattern, built column by column in order."""
        ptd = self.pt
        npt = len(ptd.pts)
        # endpoint-block pairs as (row x-index, col x-index) -> pair id
        ep_pairs = {}
        for k, (a, b) in enumerate(ptd.pairs):
            ep_pairs[(int(self.pt_x[a]), int(self.pt_x[b]))] = k
        ep_by_col = {}
        for (r, c), k in ep_pairs.items():
            ep_by_col.setdefault(c, []).append(r)
        slot = 0
        self.h_border = {}     # slot -> list of descriptors (merged)
        rows_parts, cols_parts, slot_parts = [], [], []

        def add_border(sl, desc):
            self.h_border.setdefault(sl, []).append(desc)

        def explicit_column(col, row_desc):
            """row_desc: dict row -> list of descriptors. Emits sorted rows."""
            nonlocal slot
            for r in ep_by_col.get(col, []):
                row_desc.setdefault(r, []).append(("ep", ep_pairs[(r, col)]))
            for r in sorted(row_desc):
                rows_parts.append(np.array([r], dtype=np.int64))
                cols_parts.append(np.array([col], dtype=np.int64))
                slot_parts.append(np.array([slot], dtype=np.int64))
                for d in row_desc[r]:
                    add_border(slot, d)
                slot += 1

        def segment_column(col, segs, singles):
            """Column made of long regular runs (segs: (row0, count, key)) and
            single irregular rows (dict row -> descriptors)."""
            nonlocal slot
            for r in ep_by_col.get(col, []):
                singles.setdefault(r, []).append(("ep", ep_pairs[(r, col)]))
            items = [(r0, cnt, key) for (r0, cnt, key) in segs if cnt > 0] + \
                    [(r, 1, None) for r in singles]
            items.sort(key=lambda it: it[0])
            bases = {}
            for r0, cnt, key in items:
                rows_parts.append(r0 + np.arange(cnt, dtype=np.int64))
                cols_parts.append(np.full(cnt, col, dtype=np.int64))
                slot_parts.append(slot + np.arange(cnt, dtype=np.int64))
                if key is None:
                    for d in singles[r0]:
                        add_border(slot, d)
                else:
                    bases[key] = slot
                slot += cnt
            return bases

        for ip, (ph, pd, t) in enumerate(zip(self.ir.phases, self.pd, self.ph)):
            NY, NV, N = pd.NY, pd.NV, t.N
            if N < 3:
                raise ValueError("a phase needs at least 3 mesh nodes")
            mesh = self.meshes[ip]
            # family activity at the two end nodes
            act = []
            for end in (0, 1):
                fam = {"p"}
                if end == 0:
                    n_k = int(mesh.N_K[0])
                    if np.any(self.a_nz[n_k][:, 0]):
                        fam.add("d")
                    if self._wq_nonzero(0, n_k, 0, False):
                        fam.add("i")
                else:
                    n_k = int(mesh.N_K[-1])
                    if np.any(self.a_nz[n_k][:, n_k - 1]):
                        fam.add("d")
                    if self.w_nz[n_k][n_k - 1]:
                        fam.add("i")
                act.append(fam)
            t.end_active = act
            # interior nodes are assumed fully active (true for Lobatto/Radau)
            pairs_by_b = {}
            for k, (a, b) in enumerate(pd.h2vv):
                pairs_by_b.setdefault(b, []).append((a, k))
            t.h_reg_base = np.zeros(NV, dtype=np.int64)
            t.h_nA = np.zeros(NV, dtype=np.int32)
            t.h_pair_pos = np.zeros(len(pd.h2vv), dtype=np.int32)
            for b in range(NV):
                plist = sorted(pairs_by_b.get(b, []))
                t.h_nA[b] = len(plist)
                for pos, (a, k) in enumerate(plist):
                    t.h_pair_pos[k] = pos
                for end, m in ((0, 0), (1, N - 1)):
                    col = t.x_off + b * N + m
                    if end == 1:
                        # regular run of columns (b, 1..N-2) sits before (b, N-1)
                        t.h_reg_base[b] = slot
                        nreg = N - 2
                        if plist:
                            a_arr = np.array([a for a, _ in plist], dtype=np.int64)
                            mm = np.arange(1, N - 1, dtype=np.int64)
                            rows_parts.append(
                                (t.x_off + a_arr[None, :] * N + mm[:, None]).ravel())
                            cols_parts.append(np.repeat(t.x_off + b * N + mm,
                                                        len(plist)))
                            slot_parts.append(slot + np.arange(nreg * len(plist),
                                                               dtype=np.int64))
                            slot += nreg * len(plist)
                    rd = {}
                    for a, k in plist:
                        if pd.pair_families(a, b) & act[end]:
                            rd.setdefault(t.x_off + a * N + m, []).append(
                                ("irr_vv", ip, end, k))
                    explicit_column(col, rd)
            # q columns: endpoint block only
            for i in range(ph.n_q):
                explicit_column(t.q_col + i, {})
            # free-time columns
            for tk in range(2):
                col = t.t_cols[tk]
                if col < 0:
                    continue
                segs, singles = [], {}
                for ia, a in enumerate(pd.htv):
                    segs.append((t.x_off + a * N + 1, N - 2, (tk, a)))
                    for end, m in ((0, 0), (1, N - 1)):
                        if pd.d1_families(a) & {"d", "i"} & act[end]:
                            singles.setdefault(t.x_off + a * N + m, []).append(
                                ("irr_t", ip, end, ia, tk))
                bases = segment_column(col, segs, singles)
                for key, b0 in bases.items():
                    t.h_t_base[key] = b0
        # static-parameter columns
        for j in range(self.NS):
            col = self.s_off + j
            segs, singles = [], {}
            for ip, (ph, pd, t) in enumerate(zip(self.ir.phases, self.pd, self.ph)):
                N = t.N
                for k, (a, jj) in enumerate(pd.h2vs):
                    if jj != j:
                        continue
                    segs.append((t.x_off + a * N + 1, N - 2, (ip, a, j)))
                    for end, m in ((0, 0), (1, N - 1)):
                        if pd.pair_families(a, pd.NV + j) & t.end_active[end]:
                            singles.setdefault(t.x_off + a * N + m, []).append(
                                ("irr_vs", ip, end, k))
                if j in pd.hts:
                    for tk in range(2):
                        if t.t_cols[tk] >= 0:
                            singles.setdefault(t.t_cols[tk], []).append(
                                ("red_ts", ip, pd.hts.index(j), tk))
                for k, (i, jj) in enumerate(pd.h2ss):
                    if jj == j:
                        singles.setdefault(self.s_off + i, []).append(
                            ("red_ss", ip, k))
            bases = segment_column(col, segs, singles)
            for (ip, a, jj), b0 in bases.items():
                self.ph[ip].h_s_base[(a, jj)] = b0
        self.nnz_h = slot
        self._h_parts = (rows_parts, cols_parts, slot_parts)
        self._H_rows = self._H_cols = None

    def H_structure(self):
        if self._H_rows is None:
            rp, cp, sp = self._h_parts
            rows = np.empty(self.nnz_h, dtype=np.int64)
            cols = np.empty(self.nnz_h, dtype=np.int64)
            if self.nnz_h:
                sl = np.concatenate(sp)
                assert len(sl) == self.nnz_h
                rows[sl] = np.concatenate(rp)
                cols[sl] = np.concatenate(cp)
            self._H_rows, self._H_cols = rows, cols
            self._h_parts = None
        return self._H_rows, self._H_cols

    # -------------------------------------------------------------- tiles --
    def bytes_per_node(self, pd):
        """Shared-memory doubles staged per node (see csrc/pcx_kernels.cuh)."""
        from .codegen import stage_h_rule
        hp = (len(pd.h2vv) | 1) if stage_h_rule(len(pd.h2vv)) else 0
        return 8 * (pd.NY + len(pd.d1v) + 1 + sum(1 for e, _ in pd.d1s
                                                  if pd.fam[e] == "d") + 2 + hp)

    def _build_tiles(self, sm_count, max_tile_nodes, smem_budget, tiles_per_sm=None):
        T = self.threads
        tiles = []
        total_nodes = sum(t.N for t in self.ph)
        border_phase = int(np.argmax([t.N for t in self.ph]))
        for ip, (pd, t) in enumerate(zip(self.pd, self.ph)):
            cap = max_tile_nodes or T
            cap = min(cap, max(16, smem_budget // self.bytes_per_node(pd)))
            # number of tiles: enough that no tile exceeds cap nodes, rounded up
            # to a share of a multiple of the SM count when the mesh is large
            want = max(1, int(np.ceil((t.N - 1) / max(cap - 1, 1))))
            if total_nodes >= 32 * sm_count:
                share = t.N / total_nodes
                m = max(1, int(np.ceil(want / (sm_count * share))))
                # resident CTAs per SM: the register cap's figure, or fewer when the
                # per-node staging of a large body (Delta III: ~500 B per node) makes
                # shared memory the limit (227 KB per SM, ~8 KB of static + tables)
                res = resident_ctas(T)
                stage = self.bytes_per_node(pd) * cap
                nout = (len(pd.fns) + len(pd.d1v) + len(pd.d1s) + len(pd.h2vv) + len(pd.h2vs)
                        + len(pd.h2ss) + len(pd.htv) + len(pd.hts))
                # Large bodies in the two-pass node phase (engine.smem_bytes; Delta III,
                # shuttle): a tile costs ~4 us of table prologue whatever its size and the
                # CTAs of an SM drift apart within a wave or two, so whole waves buy
                # nothing and every extra tile costs -- measured on Delta III, 10^6 nodes
                # (profiles/r02_d3_tiling.txt): 54 tiles per SM 0.338 ms against 56 at
                # 0.343; a rank of an 8-way sharded mesh 7 per SM 52.0 us, 8: 52.3, 9 (three
                # whole waves of the 3 resident CTAs): 57.7; 4-way 14: 88.9, 16: 97.4.
                # They get the fewest tiles the node cap allows.
                two_pass = nout > 40 and len(pd.h2vv) >= 16
                if stage > 40 * 1024:
                    res = max(1, min(res, (227 * 1024) // (stage + 8 * 1024)))
                if tiles_per_sm:
                    m = max(m, int(tiles_per_sm))
                elif m > res and not two_pass:
                    m = -(-m // res) * res                             # whole waves
                want = max(want, int(round(m * sm_count * share)))
            want = min(want, t.K)
            # the border pass is a CTA of its own (csrc/pcx_kernels.cuh): leave
            # it one slot of the wave so that it is resident from the start
            if (total_nodes >= 32 * sm_count and ip == border_phase and want > 2
                    and int(np.ceil((t.N - 1) / (want - 1))) + 1 <= cap):
                want -= 1
            edges = self._balanced_edges(t.sec_node, want, cap)
            for k0, k1 in zip(edges[:-1], edges[1:]):
                tiles.append((ip, int(k0), int(k1)))
        self.tile_phase = np.array([x[0] for x in tiles], dtype=np.int32)
        self.tile_k0 = np.array([x[1] for x in tiles], dtype=np.int32)
        self.tile_k1 = np.array([x[2] for x in tiles], dtype=np.int32)
        self.num_tiles = len(tiles)
        # runs of consecutive same-type sections inside each tile
        desc = np.zeros((self.num_tiles, 8), dtype=np.int64)
        run_slo, run_shi, run_type, run_gbase = [], [], [], []
        nn = []
        for it, (ip, k0, k1) in enumerate(tiles):
            t = self.ph[ip]
            nn.append(int(t.sec_node[k1] - t.sec_node[k0] + 1))
            ty = t.sec_type[k0:k1]
            cuts = np.concatenate([[0], np.flatnonzero(ty[1:] != ty[:-1]) + 1,
                                   [k1 - k0]])
            prev_rows = int(t.sec_order[k0 - 1]) - 1 if k0 > 0 else 0
            desc[it] = (ip, k0, k1, t.sec_node[k0], nn[-1], len(run_slo),
                        len(run_slo) + len(cuts) - 1, prev_rows)
            NV = t.gsec_ptr.shape[0]
            for lo, hi in zip(cuts[:-1], cuts[1:]):
                run_slo.append(int(lo))
                run_shi.append(int(hi))
                run_type.append(int(ty[lo]))
                gb = np.zeros(self.NVMAX, dtype=np.int64)
                gb[:NV] = t.gsec_ptr[:, k0 + lo]
                run_gbase.append(gb)
        self.tile_desc = desc
        self.run_slo = np.asarray(run_slo, dtype=np.int32)
        self.run_shi = np.asarray(run_shi, dtype=np.int32)
        self.run_type = np.asarray(run_type, dtype=np.int32)
        self.run_gbase = np.asarray(run_gbase, dtype=np.int64).reshape(-1, self.NVMAX)
        self.tile_nodes = np.asarray(nn, dtype=np.int32)
        self.max_tile_nodes = int(max(nn))
        self.max_tile_secs = int(np.max(self.tile_k1 - self.tile_k0))
        self.max_tile_runs = int(np.max(desc[:, 6] - desc[:, 5]))

    @staticmethod
    def _balanced_edges(sec_node, want, cap):
        """Section-range edges: ~equal node counts, never more than cap nodes."""
        K = len(sec_node) - 1
        N1 = int(sec_node[-1])
        while True:
            targets = (np.arange(1, want) * N1) / want
            cut = np.searchsorted(sec_node, targets, side="left")
            edges = np.unique(np.concatenate([[0], cut, [K]])).astype(np.int64)
            width = np.diff(sec_node[edges]) + 1
            if width.max() <= cap or want >= K:
                return edges
            want = min(K, want + max(1, want // 8))

    # ------------------------------------------------------------- border --
    def _build_border(self):
        """BV layout, reductions and the border linear map (see module doc)."""
        sidx = self.sidx
        ptd = self.pt
        # ---- BV layout ----
        bv = 1                          # [0] = 1.0
        self.red_list = []              # (ip, kind, args)
        for ip, (pd, t) in enumerate(zip(self.pd, self.ph)):
            t.red_off = bv
            t.red = {}
            for i in range(pd.NQ):
                t.red[("g", i)] = bv
                bv += 1
            for (e, j) in pd.d1s:
                if pd.fam[e] == "i":
                    t.red[("gs", e - pd.NY - pd.NP, j)] = bv
                    bv += 1
            for ij in range(len(pd.hts)):
                t.red[("hts", ij)] = bv
                bv += 1
            for k in range(len(pd.h2ss)):
                t.red[("hss", k)] = bv
                bv += 1
            t.n_red = bv - t.red_off
        self.bv_red_end = bv
        self.bv_ptval = bv              # unscaled point variable values
        bv += len(ptd.pts)
        self.bv_ptfn = bv               # J, b values
        bv += len(ptd.fns)
        self.bv_ptd1 = bv
        bv += len(ptd.d1)
        self.bv_ptd2 = bv               # contracted over (sigma*w*J, lam*W*b)
        bv += len(ptd.pairs)
        self.bv_irr0 = bv               # end-node values written by the tiles
        for ip, (pd, t) in enumerate(zip(self.pd, self.ph)):
            t.bv_irr = []
            for end in range(2):
                off = dict(vv=bv)
                bv += len(pd.h2vv)
                off["vs"] = bv
                bv += len(pd.h2vs)
                off["t"] = bv
                bv += len(pd.htv)
                t.bv_irr.append(off)
        self.bv_size = bv
        # ---- runtime scalars: [1, h_0 .. h_{P-1}] ----
        self.rs_size = 1 + self.P

        ent_grp, ent_slot, ent_ptr = [], [], [0]
        self.border_coef = Products()
        term_bv, term_rs = [], []

        def entry(grp, slot, terms):
            ent_grp.append(grp)
            ent_slot.append(slot)
            for (const, refs, bvi, rsi) in terms:
                self.border_coef.add(const, *refs)
                term_bv.append(bvi)
                term_rs.append(rsi)
            ent_ptr.append(len(term_bv))

        # ---- J and grad ----
        entry(GRP_J, 0, [(1.0, (sidx.w(),), self.bv_ptfn, 0)])
        for k, (e, a) in enumerate(ptd.d1):
            if e == 0:
                entry(GRP_GRAD, int(self.pt_x[a]),
                      [(1.0, (sidx.w(), sidx.V(self.pt_V[a])), self.bv_ptd1 + k, 0)])
        # ---- c: integral and endpoint rows ----
        pt_pos = {s: i for i, s in enumerate(ptd.pts)}
        for ip, (ph, pd, t) in enumerate(zip(self.ir.phases, self.pd, self.ph)):
            for i in range(pd.NQ):
                row = t.c_off + pd.NY * (t.N - 1) + pd.NP * t.N + i
                Wi = sidx.W(t.W_off + pd.NY + pd.NP + i)
                entry(GRP_C, row, [
                    (1.0, (Wi,), self.bv_ptval + pt_pos[ph.q[i]], 0),
                    (-1.0, (Wi,), t.red[("g", i)], 1 + ip)])
        for k in range(self.NB):
            entry(GRP_C, self.b_off + k,
                  [(1.0, (sidx.W(self.Wb_off + k),), self.bv_ptfn + 1 + k, 0)])
        # ---- G border ----
        d1_index = {ea: k for k, ea in enumerate(ptd.d1)}
        for slot, desc in self.g_border:
            kind = desc[0]
            if kind == "const":
                entry(GRP_G, slot, [(1.0, (desc[1], desc[2]), 0, 0)])
            elif kind == "b_d1":
                kb, col = desc[1], desc[2]
                a = int(np.flatnonzero(self.pt_x == col)[0])
                entry(GRP_G, slot, [(1.0, (sidx.W(self.Wb_off + kb),
                                           sidx.V(self.pt_V[a])),
                                     self.bv_ptd1 + d1_index[(kb + 1, a)], 0)])
            elif kind == "int_t":
                _, ip, i, tk = desc
                t, pd = self.ph[ip], self.pd[ip]
                Wi = sidx.W(t.W_off + pd.NY + pd.NP + i)
                sign = 1.0 if tk == 0 else -1.0     # -(+-1/2) V_t W_i (Wq.g)
                entry(GRP_G, slot, [(0.5 * sign, (Wi, sidx.V(t.t_ocp[tk])),
                                     t.red[("g", i)], 0)])
            elif kind == "int_s":
                _, ip, i, j = desc
                t, pd = self.ph[ip], self.pd[ip]
                Wi = sidx.W(t.W_off + pd.NY + pd.NP + i)
                entry(GRP_G, slot, [(-1.0, (Wi, sidx.V(self.s_ocp_off + j)),
                                     t.red[("gs", i, j)], 1 + ip)])
            else:
                raise AssertionError(kind)
        # ---- H border ----
        for slot in sorted(self.h_border):
            terms = []
            for desc in self.h_border[slot]:
                kind = desc[0]
                if kind == "ep":
                    a, b = ptd.pairs[desc[1]]
                    terms.append((1.0, (sidx.V(self.pt_V[a]), sidx.V(self.pt_V[b])),
                                  self.bv_ptd2 + desc[1], 0))
                elif kind == "irr_vv":
                    _, ip, end, k = desc
                    t, pd = self.ph[ip], self.pd[ip]
                    a, b = pd.h2vv[k]
                    terms.append((1.0, (sidx.V(t.V_off["y"] + a),
                                        sidx.V(t.V_off["y"] + b)),
                                  t.bv_irr[end]["vv"] + k, 0))
                elif kind == "irr_vs":
                    _, ip, end, k = desc
                    t, pd = self.ph[ip], self.pd[ip]
                    a, j = pd.h2vs[k]
                    terms.append((1.0, (sidx.V(t.V_off["y"] + a),
                                        sidx.V(self.s_ocp_off + j)),
                                  t.bv_irr[end]["vs"] + k, 0))
                elif kind == "irr_t":
                    _, ip, end, ia, tk = desc
                    t, pd = self.ph[ip], self.pd[ip]
                    a = pd.htv[ia]
                    sign = -0.5 if tk == 0 else 0.5
                    terms.append((sign, (sidx.V(t.V_off["y"] + a),
                                         sidx.V(t.t_ocp[tk])),
                                  t.bv_irr[end]["t"] + ia, 0))
                elif kind == "red_ts":
                    _, ip, ij, tk = desc
                    t, pd = self.ph[ip], self.pd[ip]
                    sign = -0.5 if tk == 0 else 0.5
                    terms.append((sign, (sidx.V(t.t_ocp[tk]),
                                         sidx.V(self.s_ocp_off + pd.hts[ij])),
                                  t.red[("hts", ij)], 0))
                elif kind == "red_ss":
                    _, ip, k = desc
                    t, pd = self.ph[ip], self.pd[ip]
                    i, j = pd.h2ss[k]
                    terms.append((1.0, (sidx.V(self.s_ocp_off + i),
                                        sidx.V(self.s_ocp_off + j)),
                                  t.red[("hss", k)], 0))
                else:
                    raise AssertionError(kind)
            entry(GRP_H, slot, terms)
        self.border_grp = np.asarray(ent_grp, dtype=np.int32)
        self.border_slot = np.asarray(ent_slot, dtype=np.int64)
        self.border_ptr = np.asarray(ent_ptr, dtype=np.int32)
        self.border_bv = np.asarray(term_bv, dtype=np.int32)
        self.border_rs = np.asarray(term_rs, dtype=np.int32)

"""ctypes binding of ``libpcx.so`` and the table marshalling around it.

Python is plumbing here: it builds the integer tables (``structure.py``), the
generated header (``codegen.py``) and the small scaling tables, hands them to
``pcx_create`` / ``pcx_set_scaling`` (``include/pcx.h``) and then only passes
pointers.  There is no CPU fallback: if the library cannot be loaded or the
device is absent, evaluation raises.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build as _build

PCX_HOST, PCX_DEVICE = 0, 1
EVAL_C, EVAL_DY, EVAL_JAC, EVAL_HESS, EVAL_F, EVAL_GRAD = 1, 2, 4, 8, 16, 32
EVAL_INDEPENDENT = 256        # pcx_eval_many: the argument sets are independent
EVAL_CONST_RESIDENT = 512     # PCX_HOST: the jac array keeps its iterate-independent slots

_LIB = None


class PcxError(RuntimeError):
    pass


class _Table(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("data", ctypes.c_void_p),
                ("bytes", ctypes.c_int64)]


class _Spec(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int32), ("threads", ctypes.c_int32),
                ("batch", ctypes.c_int32), ("num_tiles", ctypes.c_int32),
                ("nvmax", ctypes.c_int32), ("n_border", ctypes.c_int32),
                ("bv_size", ctypes.c_int32), ("nred_max", ctypes.c_int32),
                ("btab_len", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("num_x", ctypes.c_int64), ("num_c", ctypes.c_int64),
                ("num_dy", ctypes.c_int64), ("nnz_g", ctypes.c_int64),
                ("nnz_h", ctypes.c_int64), ("smem_bytes", ctypes.c_int64),
                ("problem_header", ctypes.c_char_p),
                ("num_tables", ctypes.c_int32),
                ("tables", ctypes.POINTER(_Table))]


class _Args(ctypes.Structure):
    """``pcx_args``: one argument set of ``pcx_eval_many``."""
    _fields_ = [(k, ctypes.c_void_p) for k in
                ("x", "lam", "sigma", "f", "grad", "c", "dy", "jac", "hess")]


def load_library(rebuild=False):
    """Load (building first if stale and nvcc is present) the C-ABI library."""
    global _LIB
    if _LIB is not None and not rebuild:
        return _LIB
    path = _build.LIB
    try:
        path = _build.build(force=rebuild)
    except Exception as exc:                      # no nvcc: use the prebuilt .so
        if not os.path.exists(path):
            raise PcxError(
                f"libpcx.so is missing and could not be built ({exc}); "
                f"pycollo_b200 has no CPU fallback") from exc
    lib = ctypes.CDLL(path)
    vp, dp, i64, i32 = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    lib.pcx_version.restype = ctypes.c_char_p
    lib.pcx_last_error.restype = ctypes.c_char_p
    lib.pcx_last_error.argtypes = [vp]
    lib.pcx_table_name.restype = ctypes.c_char_p
    lib.pcx_table_name.argtypes = [i32]
    lib.pcx_table_elem_size.argtypes = [i32]
    lib.pcx_create.argtypes = [ctypes.POINTER(_Spec), ctypes.POINTER(vp)]
    lib.pcx_create_from_file.argtypes = [ctypes.c_char_p, i32, ctypes.POINTER(vp)]
    lib.pcx_destroy.argtypes = [vp]
    lib.pcx_destroy.restype = None
    lib.pcx_set_scaling.argtypes = [vp, dp, i64, dp, i64, dp, i64, dp, i64]
    lib.pcx_eval.argtypes = [vp, i32, dp, dp, dp, dp, dp, dp, dp, dp, dp, i32, vp]
    lib.pcx_eval_f.argtypes = [vp, dp, dp, i32, vp]
    lib.pcx_eval_grad.argtypes = [vp, dp, dp, i32, vp]
    lib.pcx_eval_c.argtypes = [vp, dp, dp, i32, vp]
    lib.pcx_eval_dy.argtypes = [vp, dp, dp, i32, vp]
    lib.pcx_eval_jac.argtypes = [vp, dp, dp, i32, vp]
    lib.pcx_eval_hess.argtypes = [vp, dp, dp, dp, dp, i32, vp]
    lib.pcx_eval_jac_hess.argtypes = [vp, dp, dp, dp, dp, dp, i32, vp]
    lib.pcx_sizes.argtypes = [vp] + [ctypes.POINTER(i64)] * 5 + [ctypes.POINTER(ctypes.c_int32)]
    lib.pcx_gather.argtypes = [vp, dp, dp, i64, dp, i32, vp]
    lib.pcx_host_alloc.argtypes = [ctypes.POINTER(vp), i64]
    lib.pcx_host_free.argtypes = [vp]
    lib.pcx_host_register.argtypes = [vp, i64]
    lib.pcx_host_unregister.argtypes = [vp]
    lib.pcx_launch_count.argtypes = [vp]
    lib.pcx_launch_count.restype = i64
    lib.pcx_last_d2h_bytes.argtypes = [vp]
    lib.pcx_last_d2h_bytes.restype = i64
    lib.pcx_synchronize.argtypes = [vp, vp]
    lib.pcx_flush_l2.argtypes = [vp, i64, vp]
    lib.pcx_set_shard.argtypes = [vp, i32, i32]
    lib.pcx_exchange_alloc.argtypes = [vp, i32, ctypes.c_char_p]
    lib.pcx_exchange_attach.argtypes = [vp, i32, i32, i32, ctypes.c_char_p]
    lib.pcx_exchange_attach_ptr.argtypes = [vp, i32, i32, i32, vp]
    lib.pcx_exchange_buffer.argtypes = [vp, ctypes.POINTER(vp)]
    lib.pcx_shard_buffer.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(i64)]
    lib.pcx_apply_border.argtypes = [vp, i32, dp, dp, dp, dp, dp, dp, dp, dp, vp]
    lib.pcx_interp_guess.argtypes = [vp, dp, dp, dp, dp, dp, i32, vp]
    lib.pcx_refit_to_ph.argtypes = [vp, dp, dp, dp, i32, vp]
    lib.pcx_refit_size.argtypes = [vp, ctypes.POINTER(i64)]
    lib.pcx_mesh_error.argtypes = [vp, dp, dp, dp, dp, i32, vp]
    lib.pcx_mesh_error_sizes.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i64)]
    lib.pcx_structure_jac.argtypes = [vp, i32, dp, dp, dp]
    lib.pcx_structure_hess.argtypes = [vp, i32, dp, dp, dp]
    lib.pcx_jac_row_norms.argtypes = [vp, dp, dp, i32, vp]
    lib.pcx_expand_bounds.argtypes = [vp] + [dp] * 11 + [i32, vp]
    lib.pcx_eval_many.argtypes = [vp, i32, ctypes.POINTER(_Args), i32, i32, i32, vp, i32,
                                  ctypes.POINTER(ctypes.c_float)]
    lib.pcx_sweep_host.argtypes = [vp, i32, ctypes.POINTER(_Args), i32, i32, vp]
    lib.pcx_status.argtypes = [vp, ctypes.POINTER(ctypes.c_int)]
    lib.pcx_variant_info.argtypes = [vp, i32] + [ctypes.POINTER(ctypes.c_int)] * 4
    _LIB = lib
    return lib


_TABLE_DTYPES = {
    "tile_desc": np.int64, "run_slo": np.int32, "run_shi": np.int32,
    "run_type": np.int32, "run_gbase": np.int64, "sec_node": np.int64,
    "sec_order": np.int32, "sec_h": np.float64, "sec_type": np.int32,
    "recipes": np.uint64, "type_var_off": np.int32,
    "btab": np.float64, "order_a_off": np.int32, "order_w_off": np.int32,
    "pbase": np.int64, "border_grp": np.int32, "border_slot": np.int64,
    "border_ptr": np.int32, "border_bv": np.int32, "border_rs": np.int32,
    "pt_x": np.int64, "err_desc": np.int64,
    "refit_desc": np.int64, "refit_const": np.float64, "refit_tab": np.float64,
    "refit_off": np.int32,
}


def build_tables(S, layouts):
    """All integer / quadrature tables of one (problem, mesh) for pcx_create."""
    t = {}
    t["tile_desc"] = S.tile_desc.ravel()
    t["run_slo"], t["run_shi"], t["run_type"] = S.run_slo, S.run_shi, S.run_type
    t["run_gbase"] = S.run_gbase.ravel()
    t["sec_node"] = np.concatenate([ph.sec_node for ph in S.ph])
    t["sec_order"] = np.concatenate([ph.sec_order for ph in S.ph])
    t["sec_h"] = np.concatenate([ph.sec_h for ph in S.ph])
    t["sec_type"] = np.concatenate([ph.sec_type for ph in S.ph])
    t["recipes"] = S.recipe_words
    t["type_var_off"] = S.type_var_off.ravel()
    t["btab"] = S.btab
    t["order_a_off"], t["order_w_off"] = S.order_a_off, S.order_w_off
    pbase = []
    gsec_off = 0
    tile0 = 0
    for ip, (ph, lay) in enumerate(zip(S.ph, layouts)):
        pd = lay.pd
        pb = np.full(lay.pbase_size, -1, dtype=np.int64)
        o = lay.pb
        ntile = int(np.sum(S.tile_phase == ip))
        pb[o["N"]], pb[o["K"]] = ph.N, ph.K
        pb[o["XOFF"]], pb[o["COFF"]], pb[o["DYOFF"]] = ph.x_off, ph.c_off, ph.dy_off
        pb[o["SECOFF"]], pb[o["GSECOFF"]] = ph.sec_off, gsec_off
        pb[o["T0X"]], pb[o["TFX"]] = ph.t_cols[0], ph.t_cols[1]
        pb[o["TILE0"]], pb[o["TILE1"]] = tile0, tile0 + ntile
        pb[o["IRR0"]], pb[o["IRR1"]] = ph.bv_irr[0]["vv"], ph.bv_irr[1]["vv"]
        for i in range(pd.NY):
            pb[o["GT0"] + i] = ph.g_tcol_base.get((0, i), -1)
            pb[o["GTF"] + i] = ph.g_tcol_base.get((1, i), -1)
        for k, (e, j) in enumerate(pd.d1s):
            pb[o["GSCOL"] + k] = ph.g_scol_base.get((j, e), -1)
        for b in range(pd.NV):
            pb[o["HREG"] + b] = ph.h_reg_base[b]
        for k, (a, j) in enumerate(pd.h2vs):
            pb[o["HS"] + k] = ph.h_s_base.get((a, j), -1)
        for k, a in enumerate(pd.htv):
            pb[o["HT0"] + k] = ph.h_t_base.get((0, a), -1)
            pb[o["HTF"] + k] = ph.h_t_base.get((1, a), -1)
        pbase.append(pb)
        gsec_off += ph.gsec_ptr.size
        tile0 += ntile
    t["pbase"] = np.concatenate(pbase)
    t["border_grp"], t["border_slot"] = S.border_grp, S.border_slot
    t["border_ptr"], t["border_bv"], t["border_rs"] = S.border_ptr, S.border_bv, S.border_rs
    t["pt_x"] = S.pt_x
    # mesh-error pass (pcx_mesh_error): per phase x_off, c_off, N, K, n_y, offset
    # of the phase's K+1 section-start nodes in sec_node, offsets of its block in
    # the error arrays / per-section maxima, and mmax = max_k N_k - 1
    ed = np.zeros((len(S.ph), 12), dtype=np.int64)
    eo = so = 0
    for ip, (ph, lay) in enumerate(zip(S.ph, layouts)):
        mmax = int(np.max(ph.sec_order)) - 1
        ed[ip, :9] = (ph.x_off, ph.c_off, ph.N, ph.K, lay.pd.NY, ph.sec_off + ip,
                      eo, so, mmax)
        eo += ph.K * lay.pd.NY * mmax
        so += ph.K
    t["err_desc"] = ed.ravel()
    # solution re-fit onto the p+1 mesh (pcx_refit_to_ph): per-order Cy / Pu
    # matrices (quadrature.refit_matrices) and the per-phase layout of x and x_ph
    quad = S.meshes[0].quadrature
    omax = max(S.orders)
    tab, off = [], np.zeros(2 * (omax + 1), dtype=np.int32)
    for n in S.orders:
        Cy, Pu = quad.refit_matrices(n)
        off[2 * n] = sum(len(a) for a in tab)
        tab.append(Cy.ravel())
        off[2 * n + 1] = sum(len(a) for a in tab)
        tab.append(Pu.ravel())
    t["refit_tab"] = np.concatenate(tab)
    t["refit_off"] = off
    rd = np.zeros((len(S.ph) + 1, 16), dtype=np.int64)
    rc = np.zeros((len(S.ph), 2))
    xph = work = 0
    for ip, (ph, lay, irp) in enumerate(zip(S.ph, layouts, S.ir.phases)):
        pd = lay.pd
        NU = pd.NV - pd.NY
        nqt = irp.n_q + irp.n_t
        rd[ip, :12] = (ph.x_off, ph.N, ph.K, pd.NY, NU, nqt, ph.dy_off, ph.sec_off + ip,
                       ph.sec_off, xph, ph.t_cols[0], ph.t_cols[1])
        rd[ip, 12:15] = (ph.c_off, irp.n_p, irp.n_q)            # pcx_expand_bounds
        rc[ip] = ph.t_const
        xph += pd.NV * (ph.N + ph.K) + nqt
        work += ph.K * pd.NV
    rd[-1, :5] = (S.s_off, S.NS, xph, xph + S.NS, work)
    rd[-1, 5:7] = (S.NB, S.b_off)
    t["refit_desc"] = rd.ravel()
    t["refit_const"] = rc.ravel()
    return {k: np.ascontiguousarray(v, dtype=_TABLE_DTYPES[k]) for k, v in t.items()}


def scaling_tables(S, layouts, V_ocp, r_ocp, W_ocp, w):
    """pscal / gscal / border_coef / pt_scal for pcx_set_scaling.

    Reference: ``x = V*x_tilde + r`` (``pycollo/scaling.py:172-178``), constraint
    scaling W per OCP-level constraint and objective scaling w
    (``pycollo/backend.py:1465-1493``, ``scaling.py:346-430``)."""
    V_ocp = np.asarray(V_ocp, dtype=np.float64)
    r_ocp = np.asarray(r_ocp, dtype=np.float64)
    W_ocp = np.asarray(W_ocp, dtype=np.float64)
    ps_all = []
    Vs = V_ocp[S.s_ocp_off:S.s_ocp_off + S.NS]
    rs = r_ocp[S.s_ocp_off:S.s_ocp_off + S.NS]
    for ph, lay, irp in zip(S.ph, layouts, S.ir.phases):
        pd = lay.pd
        ps = np.zeros(lay.pscal_size)
        o = lay.ps
        Vv = V_ocp[ph.V_off["y"]:ph.V_off["y"] + pd.NV]
        rv = r_ocp[ph.V_off["y"]:ph.V_off["y"] + pd.NV]
        Wf = W_ocp[ph.W_off:ph.W_off + pd.NF]
        ps[o["VV"]:o["VV"] + pd.NV] = Vv
        ps[o["RV"]:o["RV"] + pd.NV] = rv
        ps[o["WFN"]:o["WFN"] + pd.NF] = Wf
        for k, (e, a) in enumerate(pd.d1v):
            ps[o["D1V"] + k] = Wf[e] * Vv[a]
        for k, (e, j) in enumerate(pd.d1s):
            ps[o["D1S"] + k] = Wf[e] * Vs[j]
        for i in range(pd.NY):
            ps[o["GCST"] + 1 + 2 * i] = Wf[i] * Vv[i]
            ps[o["GCST"] + 2 + 2 * i] = -(Wf[i] * Vv[i])
        for k, (a, b) in enumerate(pd.h2vv):
            ps[o["H2VV"] + k] = Vv[a] * Vv[b]
        for k, (a, j) in enumerate(pd.h2vs):
            ps[o["H2VS"] + k] = Vv[a] * Vs[j]
        Vt = [V_ocp[i] if i >= 0 else 0.0 for i in ph.t_ocp]
        rt = [r_ocp[i] if i >= 0 else c for i, c in zip(ph.t_ocp, ph.t_const)]
        for k, a in enumerate(pd.htv):
            ps[o["HT0"] + k] = -0.5 * Vt[0] * Vv[a]
            ps[o["HTF"] + k] = 0.5 * Vt[1] * Vv[a]
        for i in range(pd.NY):
            ps[o["GT0"] + i] = -0.5 * Vt[0] * Wf[i]
            ps[o["GTF"] + i] = 0.5 * Vt[1] * Wf[i]
        ps[o["TINFO"]:o["TINFO"] + 4] = [Vt[0], rt[0], Vt[1], rt[1]]
        # literals of a shared expression body that differ between its phases (codegen.share_groups)
        ps[o["KC"]:o["KC"] + len(lay.kc)] = lay.kc
        ps_all.append(ps)
    pscal = np.concatenate(ps_all)
    gscal = np.concatenate([Vs, rs, [float(w)],
                            W_ocp[S.Wb_off:S.Wb_off + S.NB]]).astype(np.float64)
    scales = S.sidx.vector(V_ocp, W_ocp, float(w))
    coef = S.border_coef.evaluate(scales)
    pt_scal = np.concatenate([V_ocp[S.pt_V], r_ocp[S.pt_V]]).astype(np.float64)
    return pscal, gscal, coef, pt_scal


def smem_bytes(S, layouts, threads):
    NN, SS = S.max_tile_nodes, S.max_tile_secs
    best = 0
    for lay in layouts:
        pd = lay.pd
        nds = sum(1 for e, _ in pd.d1s if pd.fam[e] == "d")
        from .codegen import stage_h_rule
        hp = (len(pd.h2vv) | 1) if stage_h_rule(len(pd.h2vv)) else 0
        # two-pass node phase (large bodies): the second pass stages NH2VV * threads
        # Hessian entries over sD / sDP / sDS / sLam; room beyond them only if they are smaller
        nout = (len(pd.fns) + len(pd.d1v) + len(pd.d1s) + len(pd.h2vv) + len(pd.h2vs)
                + len(pd.h2ss) + len(pd.htv) + len(pd.hts))
        two_extra = 0
        if nout > 40 and len(pd.h2vv) >= 16 and not hp:             # PCX_TWO_PASS_MIN
            # (the Hessian-only variant stages them NEXT to the multipliers: the room is
            # what the first-derivative staging of the fused variant occupies)
            region = (len(pd.d1v) + nds) * (NN | 1) + len(pd.d1v) * (SS + 1)
            two_extra = max(0, len(pd.h2vv) * threads - region)
        dbl = (len(S.btab) + SS + 1 + (pd.NY + len(pd.d1v) + nds) * (NN | 1)
               + len(pd.d1v) * (SS + 1)
               + pd.NY * (NN + 24) + hp * NN + two_extra + threads // 32 + 2)      # PCX_LAM_HALO
        ints = (SS + 2) + (SS + 1) + NN + 2 * (SS + 1) + 8
        best = max(best, 8 * dbl + 4 * ints)
    best = max(best, 8 * (32 + S.bv_size))      # border pass: scratch + BV
    return int((best + 15) // 16 * 16)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(ctypes.c_void_p)
    if hasattr(a, "data_ptr"):                  # torch tensor
        return ctypes.c_void_p(a.data_ptr())
    return ctypes.c_void_p(int(a))


class Engine:
    """One compiled problem on one mesh on one device (wraps ``pcx_engine``)."""

    # above this many Jacobian non-zeros the int64 patterns (16 B per entry, host
    # copies inside the library) are only handed over on request
    STRUCTURE_AUTO_LIMIT = 50_000_000

    def __init__(self, S, layouts, header, *, batch=1, device=0, min_blocks=None,
                 structure=None):
        spec = self._prepare(S, layouts, header, batch, device, min_blocks, structure)
        h = ctypes.c_void_p()
        rc = self.lib.pcx_create(ctypes.byref(spec), ctypes.byref(h))
        if rc != 0:
            raise PcxError(f"pcx_create failed ({rc}): "
                           f"{self.lib.pcx_last_error(None).decode(errors='replace')}")
        self.h = h

    @classmethod
    def write_spec(cls, path, S, layouts, header, scaling=None, *, batch=1, min_blocks=None,
                   structure=None):
        """``save_spec`` without creating an engine (no device needed): the file a
        compiled host hands to ``pcx_create_from_file``.  ``scaling`` =
        ``(V_ocp, r_ocp, W_ocp, w)`` or None."""
        self = cls.__new__(cls)
        self.h = None
        self._prepare(S, layouts, header, batch, 0, min_blocks, structure)
        if scaling is not None:
            self._scal = scaling_tables(S, layouts, *scaling)
        return self.save_spec(path)

    def _prepare(self, S, layouts, header, batch, device, min_blocks, structure):
        self.lib = load_library()
        self.S, self.layouts = S, layouts
        self.batch = int(batch)
        self.device = int(device)
        self.threads = int(S.threads)
        self.tables = build_tables(S, layouts)
        if S.max_tile_nodes > self.threads:
            raise PcxError(f"a tile holds {S.max_tile_nodes} nodes but the CTA has only "
                           f"{self.threads} threads (one thread per node)")
        self.smem = smem_bytes(S, layouts, self.threads)
        if self.smem > 227 * 1024:
            raise PcxError(f"tile needs {self.smem} B of shared memory (> 227 KB)")
        n = self.lib.pcx_table_count()
        if structure is None:
            structure = S.nnz_g <= self.STRUCTURE_AUTO_LIMIT
        self.has_structure = bool(structure)
        extra = []
        if structure:       # patterns for pcx_structure_* / pcx_jac_row_norms (host side)
            gr, gc = S.G_structure()
            hr, hc = S.H_structure()
            extra = [(b"g_rows", gr), (b"g_cols", gc), (b"h_rows", hr), (b"h_cols", hc)]
            cr = S.G_constant_ranges()
            if len(cr):
                extra.append((b"g_const_ranges", cr.ravel()))
        arr = (_Table * (n + len(extra)))()
        self._keep, self._table_names = [], []
        for i in range(n):
            name = self.lib.pcx_table_name(i)
            a = self.tables[name.decode()]
            assert a.itemsize == self.lib.pcx_table_elem_size(i), name
            self._keep.append(a)
            self._table_names.append(name)
            arr[i] = _Table(name, a.ctypes.data_as(ctypes.c_void_p), a.nbytes)
        for k, (name, a) in enumerate(extra):
            a = np.ascontiguousarray(a, dtype=np.int64)
            self._keep.append(a)
            self._table_names.append(name)
            arr[n + k] = _Table(name, a.ctypes.data_as(ctypes.c_void_p), a.nbytes)
        n += len(extra)
        self._header = header.encode()
        self._spec_fields = dict(
            threads=self.threads, batch=self.batch, num_tiles=S.num_tiles, nvmax=S.NVMAX,
            n_border=len(S.border_grp), bv_size=S.bv_size,
            nred_max=max([l.nred for l in layouts] + [1]), btab_len=len(S.btab),
            reserved=self._min_blocks(min_blocks, S),
            num_x=S.num_x, num_c=S.num_c, num_dy=S.num_dy,
            nnz_g=S.nnz_g, nnz_h=S.nnz_h, smem_bytes=self.smem)
        self._tables_arr = arr
        return _Spec(device=self.device, problem_header=self._header, num_tables=n,
                     tables=arr, **self._spec_fields)

    @staticmethod
    def _min_blocks(min_blocks, S):
        """CTAs to keep resident per SM: one full wave on large meshes."""
        if min_blocks is None and "PCX_MIN_BLOCKS" in os.environ:
            min_blocks = int(os.environ["PCX_MIN_BLOCKS"])
        if min_blocks is None:
            from .structure import resident_ctas
            per_sm = -(-S.num_tiles * 1 // 148)
            min_blocks = max(1, min(resident_ctas(S.threads), per_sm))
        return int(min_blocks)

    def __del__(self):
        h = getattr(self, "h", None)
        if h:
            self.lib.pcx_destroy(h)
            self.h = None

    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.pcx_last_error(self.h).decode(errors="replace")
            raise PcxError(f"{what} failed ({rc}): {msg}")

    def set_scaling(self, V_ocp, r_ocp, W_ocp, w):
        ps, gs, bc, pt = scaling_tables(self.S, self.layouts, V_ocp, r_ocp, W_ocp, w)
        self._scal = (ps, gs, bc, pt)
        self._check(self.lib.pcx_set_scaling(
            self.h, _ptr(ps), ps.size, _ptr(gs), gs.size, _ptr(bc), bc.size,
            _ptr(pt), pt.size), "pcx_set_scaling")

    def save_spec(self, path):
        """Write everything ``pcx_create`` and ``pcx_set_scaling`` were given into one
        file a compiled host opens with ``pcx_create_from_file`` (``include/pcx.h``
        documents the layout): the symbolic problem definition stays in Python, as in
        the reference, but the process that SOLVES needs no Python."""
        f = self._spec_fields

        def block(b):
            return b + b"\0" * (-len(b) % 8)
        out = [b"PCXSPEC1",
               block(np.array([f["threads"], f["batch"], f["num_tiles"], f["nvmax"],
                               f["n_border"], f["bv_size"], f["nred_max"], f["btab_len"],
                               f["reserved"], len(self._keep)], dtype="<i4").tobytes()),
               np.array([f["num_x"], f["num_c"], f["num_dy"], f["nnz_g"], f["nnz_h"],
                         f["smem_bytes"]], dtype="<i8").tobytes()]
        text = self._header + b"\0"
        out += [np.int64(len(text)).tobytes(), block(text)]
        for name, a in zip(self._table_names, self._keep):
            nm = name + b"\0"
            out += [np.int64(len(nm)).tobytes(), block(nm),
                    np.int64(a.nbytes).tobytes(), block(a.tobytes())]
        scal = getattr(self, "_scal", None) or [np.zeros(0)] * 4
        out.append(np.array([a.size for a in scal], dtype="<i8").tobytes())
        out += [np.ascontiguousarray(a, dtype="<f8").tobytes() for a in scal]
        with open(path, "wb") as fh:
            fh.write(b"".join(out))
        return path

    @staticmethod
    def read_spec(path):
        """Parse a spec file back (the layout ``include/pcx.h`` documents): dict with the
        scalar ``fields``, the ``header`` text, ``tables`` (name -> raw bytes) and the four
        ``scaling`` arrays.  For inspection and for pinning the format in the tests; the
        reader that matters is ``pcx_create_from_file``."""
        raw = open(path, "rb").read()
        if raw[:8] != b"PCXSPEC1" or len(raw) % 8:
            raise PcxError("not a PCXSPEC1 file")
        pos = 8

        def take(n):
            nonlocal pos
            out = raw[pos:pos + n]
            if len(out) != n:
                raise PcxError("truncated spec file")
            pos += n + (-n % 8)
            return out
        i32 = np.frombuffer(take(40), dtype="<i4")
        i64 = np.frombuffer(take(48), dtype="<i8")
        names = ("threads", "batch", "num_tiles", "nvmax", "n_border", "bv_size", "nred_max",
                 "btab_len", "reserved", "num_tables", "num_x", "num_c", "num_dy", "nnz_g",
                 "nnz_h", "smem_bytes")
        fields = dict(zip(names, [int(v) for v in i32] + [int(v) for v in i64]))
        header = take(int(np.frombuffer(take(8), dtype="<i8")[0]))[:-1].decode()
        tables = {}
        for _ in range(fields["num_tables"]):
            name = take(int(np.frombuffer(take(8), dtype="<i8")[0]))[:-1].decode()
            tables[name] = take(int(np.frombuffer(take(8), dtype="<i8")[0]))
        ns = np.frombuffer(take(32), dtype="<i8")
        scaling = [np.frombuffer(take(8 * int(n)), dtype="<f8") for n in ns]
        if pos != len(raw):
            raise PcxError("trailing bytes in spec file")
        return dict(fields=fields, header=header, tables=tables, scaling=scaling)

    # -- host-space convenience (numpy in / numpy out) ---------------------
    def eval_host(self, what, x, lam=None, sigma=None):
        S, B = self.S, self.batch
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(B, S.num_x)
        lam_ = None if lam is None else \
            np.ascontiguousarray(lam, dtype=np.float64).reshape(B, S.num_c)
        sig_ = None if sigma is None else \
            np.ascontiguousarray(np.broadcast_to(np.asarray(sigma, dtype=np.float64),
                                                 (B,)))
        out = {}
        if what & EVAL_F:
            out["f"] = np.empty(B)
        if what & EVAL_GRAD:
            out["grad"] = np.empty((B, S.num_x))
        if what & EVAL_C:
            out["c"] = np.empty((B, S.num_c))
        if what & EVAL_DY:
            out["dy"] = np.empty((B, S.num_dy))
        if what & EVAL_JAC:
            out["jac"] = np.empty((B, S.nnz_g))
        if what & EVAL_HESS:
            out["hess"] = np.empty((B, S.nnz_h))
        self._check(self.lib.pcx_eval(
            self.h, what, _ptr(x), _ptr(lam_), _ptr(sig_), _ptr(out.get("f")),
            _ptr(out.get("grad")), _ptr(out.get("c")), _ptr(out.get("dy")),
            _ptr(out.get("jac")), _ptr(out.get("hess")), PCX_HOST, None), "pcx_eval")
        return out

    # -- structure / scaling / bounds through the C ABI -------------------------
    _ORDERS = {"ccs": 0, "triu_ccs": 0, "native": 0, "row_major": 1, "tril_row_major": 1}

    def _structure(self, fn, nnz, order):
        rows, cols, perm = (np.empty(nnz, dtype=np.int64) for _ in range(3))
        self._check(fn(self.h, self._ORDERS[order], _ptr(rows), _ptr(cols), _ptr(perm)),
                    fn.__name__)
        return rows, cols, perm

    def structure_jac(self, order="ccs"):
        """``pcx_structure_jac``: (rows, cols, perm); ``values[perm]`` is the order."""
        return self._structure(self.lib.pcx_structure_jac, self.S.nnz_g, order)

    def structure_hess(self, order="triu_ccs"):
        return self._structure(self.lib.pcx_structure_hess, self.S.nnz_h, order)

    def jac_row_norms_host(self, x):
        """``pcx_jac_row_norms``: ||G[i, :]||_2 per constraint row, reduced on the
        device (``pycollo/scaling.py:392-396`` without the dense matrix)."""
        S, B = self.S, self.batch
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(B, S.num_x)
        out = np.empty((B, S.num_c))
        self._check(self.lib.pcx_jac_row_norms(self.h, _ptr(x), _ptr(out), PCX_HOST, None),
                    "pcx_jac_row_norms")
        return out[0] if B == 1 else out

    def expand_bounds_host(self, ocp_x_bnd, y_t0_bnd, y_tF_bnd, ocp_c_bnd, V_ocp, r_ocp, W_ocp):
        """``pcx_expand_bounds``: scaled (x_lo, x_hi, c_lo, c_hi) on the mesh."""
        S = self.S
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        xl, xh = np.empty(S.num_x), np.empty(S.num_x)
        cl, ch = np.empty(S.num_c), np.empty(S.num_c)
        args = [f(ocp_x_bnd), f(y_t0_bnd), f(y_tF_bnd), f(ocp_c_bnd), f(V_ocp), f(r_ocp), f(W_ocp)]
        self._check(self.lib.pcx_expand_bounds(
            self.h, *[_ptr(a) if a.size else None for a in args], _ptr(xl), _ptr(xh),
            _ptr(cl), _ptr(ch), PCX_HOST, None), "pcx_expand_bounds")
        return xl, xh, cl, ch

    def make_args(self, sets):
        """Argument sets for ``eval_many`` from dicts of device tensors / pointers."""
        arr = (_Args * len(sets))()
        for i, d in enumerate(sets):
            for k in ("x", "lam", "sigma", "f", "grad", "c", "dy", "jac", "hess"):
                p = _ptr(d.get(k))
                setattr(arr[i], k, p.value if p is not None else None)
        arr.keepalive = sets
        return arr

    def eval_many(self, what, args, count, stream=None, gate=True, timed=True, warm=0):
        """``pcx_eval_many``: ``warm`` untimed + ``count`` timed launches enqueued from
        C, cycling through the argument sets; returns the device time in ms between
        the first timed launch and the end of the last one (CUDA events on
        ``stream``) when ``timed``."""
        ms = ctypes.c_float(0.0)
        self._check(self.lib.pcx_eval_many(
            self.h, what, args, len(args), int(count), int(warm),
            ctypes.c_void_p(stream) if stream else None, 1 if gate else 0,
            ctypes.byref(ms) if timed else None), "pcx_eval_many")
        return float(ms.value)

    def sweep_host(self, what, args, count, stream=None):
        """``pcx_sweep_host``: ``count`` pipelined host-space evaluations cycling through
        the (>= 2) host argument sets of ``make_args``; blocks until all are done."""
        self._check(self.lib.pcx_sweep_host(
            self.h, what, args, len(args), int(count),
            ctypes.c_void_p(stream) if stream else None), "pcx_sweep_host")

    def variant_info(self, what):
        """dict(blocks_per_sm, registers, local_bytes, static_smem_bytes) of a variant."""
        v = [ctypes.c_int(0) for _ in range(4)]
        self._check(self.lib.pcx_variant_info(self.h, what, *[ctypes.byref(x) for x in v]),
                    "pcx_variant_info")
        return dict(zip(("blocks_per_sm", "registers", "local_bytes", "static_smem_bytes"),
                        (int(x.value) for x in v)), dynamic_smem_bytes=self.smem)

    def status(self):
        code = ctypes.c_int(0)
        self._check(self.lib.pcx_status(self.h, ctypes.byref(code)), "pcx_status")
        return int(code.value)

    # -- one mesh over several GPUs (include/pcx.h) ----
    def set_shard(self, tile_begin, tile_end):
        self._check(self.lib.pcx_set_shard(self.h, int(tile_begin), int(tile_end)),
                    "pcx_set_shard")

    def exchange_alloc(self, world):
        """Border rank: allocate the peer-memory exchange buffer; returns its CUDA
        IPC handle (64 bytes) for the other ranks."""
        buf = ctypes.create_string_buffer(64)
        self._check(self.lib.pcx_exchange_alloc(self.h, int(world), buf), "pcx_exchange_alloc")
        return buf.raw

    def exchange_attach(self, rank, world, border_rank, handle=None):
        self._check(self.lib.pcx_exchange_attach(self.h, int(rank), int(world), int(border_rank),
                                                 handle), "pcx_exchange_attach")

    def exchange_attach_ptr(self, rank, world, border_rank, base):
        self._check(self.lib.pcx_exchange_attach_ptr(self.h, int(rank), int(world),
                                                     int(border_rank), ctypes.c_void_p(base)),
                    "pcx_exchange_attach_ptr")

    def exchange_buffer(self):
        ptr = ctypes.c_void_p()
        self._check(self.lib.pcx_exchange_buffer(self.h, ctypes.byref(ptr)), "pcx_exchange_buffer")
        return int(ptr.value)

    def shard_buffer(self):
        """(device pointer, length) of the small border-exchange buffer."""
        ptr, n = ctypes.c_void_p(), ctypes.c_int64()
        self._check(self.lib.pcx_shard_buffer(self.h, ctypes.byref(ptr), ctypes.byref(n)),
                    "pcx_shard_buffer")
        return int(ptr.value), int(n.value)

    def apply_border(self, what, x, lam=None, sigma=None, f=None, grad=None, c=None,
                     jac=None, hess=None, stream=None):
        self._check(self.lib.pcx_apply_border(
            self.h, what, _ptr(x), _ptr(lam), _ptr(sigma), _ptr(f), _ptr(grad), _ptr(c),
            _ptr(jac), _ptr(hess), ctypes.c_void_p(stream) if stream else None),
            "pcx_apply_border")

    # -- guess interpolation to this engine's mesh (row N3) ----------------------
    def interp_guess_host(self, x_prev, prev_tau, tau):
        """``pcx_interp_guess``: ``x_prev`` in the x layout of a previous mesh whose
        per-phase abscissae are the arrays ``prev_tau``; ``tau`` = this mesh's."""
        S = self.S
        xp = np.ascontiguousarray(x_prev, dtype=np.float64)
        pn = np.ascontiguousarray([len(t) for t in prev_tau], dtype=np.int64)
        pt = np.ascontiguousarray(np.concatenate(prev_tau), dtype=np.float64)
        tt = np.ascontiguousarray(np.concatenate(tau), dtype=np.float64)
        out = np.empty(S.num_x)
        self._check(self.lib.pcx_interp_guess(self.h, _ptr(xp), _ptr(pt), _ptr(pn), _ptr(tt),
                                              _ptr(out), PCX_HOST, None), "pcx_interp_guess")
        return out

    # -- solution re-fit onto the p+1 mesh (row N2) --------------------------
    def refit_size(self):
        n = ctypes.c_int64()
        self._check(self.lib.pcx_refit_size(self.h, ctypes.byref(n)), "pcx_refit_size")
        return int(n.value)

    def refit_to_ph_host(self, x_user, dy):
        S, B = self.S, self.batch
        x = np.ascontiguousarray(x_user, dtype=np.float64).reshape(B, S.num_x)
        d = np.ascontiguousarray(dy, dtype=np.float64).reshape(B, S.num_dy)
        out = np.empty((B, self.refit_size()))
        self._check(self.lib.pcx_refit_to_ph(self.h, _ptr(x), _ptr(d), _ptr(out), PCX_HOST, None),
                    "pcx_refit_to_ph")
        return out[0] if B == 1 else out

    def refit_to_ph_ptr(self, x_user, dy, x_ph, space=PCX_DEVICE, stream=None):
        self._check(self.lib.pcx_refit_to_ph(
            self.h, _ptr(x_user), _ptr(dy), _ptr(x_ph), space,
            ctypes.c_void_p(stream) if stream else None), "pcx_refit_to_ph")

    def mesh_error_sizes(self):
        ne, ns = ctypes.c_int64(), ctypes.c_int64()
        self._check(self.lib.pcx_mesh_error_sizes(self.h, ctypes.byref(ne), ctypes.byref(ns)),
                    "pcx_mesh_error_sizes")
        return int(ne.value), int(ns.value)

    def mesh_error_host(self, x_ph):
        """``pcx_mesh_error`` on host arrays: per phase ``(abs, rel, max_rel)`` with
        ``abs/rel`` of shape ``(K, n_y, mmax)`` and ``max_rel`` of shape ``(K,)``
        (``pycollo/mesh_refinement.py:206-240``); a leading batch axis if batch > 1."""
        S, B = self.S, self.batch
        x = np.ascontiguousarray(x_ph, dtype=np.float64).reshape(B, S.num_x)
        ne, ns = self.mesh_error_sizes()
        a, r, m = np.empty((B, ne)), np.empty((B, ne)), np.empty((B, ns))
        self._check(self.lib.pcx_mesh_error(self.h, _ptr(x), _ptr(a), _ptr(r), _ptr(m),
                                            PCX_HOST, None), "pcx_mesh_error")
        ed = self.tables["err_desc"].reshape(-1, 12)
        out = []
        for row in ed:
            K, NY, eo, so, mmax = int(row[3]), int(row[4]), int(row[6]), int(row[7]), int(row[8])
            shp = (B, K, NY, mmax)
            pa = a[:, eo:eo + K * NY * mmax].reshape(shp)
            pr = r[:, eo:eo + K * NY * mmax].reshape(shp)
            pm = m[:, so:so + K]
            out.append((pa[0], pr[0], pm[0]) if B == 1 else (pa, pr, pm))
        return out

    def mesh_error_ptr(self, x_ph, abs_err=None, rel_err=None, max_rel=None,
                       space=PCX_DEVICE, stream=None):
        self._check(self.lib.pcx_mesh_error(
            self.h, _ptr(x_ph), _ptr(abs_err), _ptr(rel_err), _ptr(max_rel), space,
            ctypes.c_void_p(stream) if stream else None), "pcx_mesh_error")

    # -- raw pointers (device tensors, pinned buffers) ------------------------
    def eval_ptr(self, what, x, lam=None, sigma=None, f=None, grad=None, c=None,
                 dy=None, jac=None, hess=None, space=PCX_DEVICE, stream=None):
        self._check(self.lib.pcx_eval(
            self.h, what, _ptr(x), _ptr(lam), _ptr(sigma), _ptr(f), _ptr(grad),
            _ptr(c), _ptr(dy), _ptr(jac), _ptr(hess), space,
            ctypes.c_void_p(stream) if stream else None), "pcx_eval")

    def bind(self, what, x, lam=None, sigma=None, f=None, grad=None, c=None, dy=None,
             jac=None, hess=None, space=PCX_DEVICE, stream=None):
        """``pcx_eval`` with every argument converted once: returns a zero-argument
        callable (one foreign call, ~2 us of host time instead of ~7 us through
        ``eval_ptr``).  The buffers must stay alive and in place."""
        args = (self.h, what, _ptr(x), _ptr(lam), _ptr(sigma), _ptr(f), _ptr(grad), _ptr(c),
                _ptr(dy), _ptr(jac), _ptr(hess), space,
                ctypes.c_void_p(stream) if stream else None)
        fn, check = self.lib.pcx_eval, self._check

        def call():
            rc = fn(*args)
            if rc:
                check(rc, "pcx_eval")
        call.keepalive = (x, lam, sigma, f, grad, c, dy, jac, hess)
        return call

    def gather(self, src, perm, n, dst, stream=None):
        self._check(self.lib.pcx_gather(self.h, _ptr(src), _ptr(perm), n, _ptr(dst),
                                        PCX_DEVICE,
                                        ctypes.c_void_p(stream) if stream else None),
                    "pcx_gather")

    def flush_l2(self, nbytes=256 << 20, stream=None):
        self._check(self.lib.pcx_flush_l2(
            self.h, nbytes, ctypes.c_void_p(stream) if stream else None),
            "pcx_flush_l2")

    def synchronize(self, stream=None):
        self._check(self.lib.pcx_synchronize(
            self.h, ctypes.c_void_p(stream) if stream else None), "pcx_synchronize")

    @property
    def last_d2h_bytes(self):
        return int(self.lib.pcx_last_d2h_bytes(self.h))

    @property
    def launch_count(self):
        return int(self.lib.pcx_launch_count(self.h))

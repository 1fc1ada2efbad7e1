"""CPU emulation of the device algorithm (TEST INFRASTRUCTURE).

Walks the *same tables* the CUDA kernels consume (``engine.build_tables``,
``engine.scaling_tables``) with the same control flow as
``csrc/pcx_kernels.cuh`` -- tiles, staged node values, recipe decoding, row
contractions, reductions, border map -- in plain Python/numpy, so that table
construction can be validated against the oracle without a GPU.  It is not a
product path and is far too slow for real meshes.
"""
from __future__ import annotations

import numpy as np
import sympy as sym

from pycollo_b200 import structure as st

F_C, F_DY, F_G, F_H, F_J, F_GRAD = 1, 2, 4, 8, 16, 32


def _node_fn(pd, NS):
    NV, NF = pd.NV, pd.NF
    v = list(pd.variables)
    muh = [sym.Symbol(f"_mh{e}") for e in range(NF)]
    mut = [sym.Symbol(f"_mt{e}") for e in range(NF)]
    contr = {}
    for (e, a, b, dab) in pd.d2:
        contr[(a, b)] = contr.get((a, b), 0) + muh[e] * dab
    d1v = {}
    for (e, a), de in zip(pd.d1v, pd.d1v_expr):
        if pd.fam[e] in "di":
            d1v[a] = d1v.get(a, 0) + mut[e] * de
    for (e, j), de in zip(pd.d1s, pd.d1s_expr):
        if pd.fam[e] in "di":
            d1v[NV + j] = d1v.get(NV + j, 0) + mut[e] * de
    outs = dict(
        F=list(pd.fns), D1V=list(pd.d1v_expr), D1S=list(pd.d1s_expr),
        H2VV=[contr[ab] for ab in pd.h2vv],
        H2VS=[contr[(a, NV + j)] for a, j in pd.h2vs],
        H2SS=[contr[(NV + i, NV + j)] for i, j in pd.h2ss],
        HTV=[d1v[a] for a in pd.htv], HTS=[d1v[NV + j] for j in pd.hts])
    keys = list(outs)
    flat = [e for k in keys for e in outs[k]]
    fn = sym.lambdify([v, muh, mut], flat, modules="math")
    sizes = [len(outs[k]) for k in keys]

    def call(vv, mh, mt):
        r = fn(list(vv), list(mh), list(mt))
        res, o = {}, 0
        for k, n in zip(keys, sizes):
            res[k] = [float(z) for z in r[o:o + n]]
            o += n
        return res
    return call


def _point_fn(ptd, NB):
    pts = list(ptd.pts)
    mult = [sym.Symbol(f"_m{e}") for e in range(1 + NB)]
    contr = {}
    for (e, a, b, dab) in ptd.d2:
        contr[(a, b)] = contr.get((a, b), 0) + mult[e] * dab
    flat = list(ptd.fns) + list(ptd.d1_expr) + [contr[ab] for ab in ptd.pairs]
    fn = sym.lambdify([pts, mult], flat, modules="math")
    n0, n1 = len(ptd.fns), len(ptd.d1_expr)

    def call(pt, m):
        r = [float(z) for z in fn(list(pt), list(m))]
        return r[:n0], r[n0:n0 + n1], r[n0 + n1:]
    return call


def emulate(low, tables, scal, x, lam=None, sigma=1.0, flags=F_G | F_H,
            tile_range=None, stage=0, xbuf=None):
    """stage 0: whole evaluation.  Sharded (pcx_set_shard / pcx_apply_border):
    stage 1 walks ``tile_range`` only and returns this rank's border share in
    ``out["xbuf"]`` instead of applying the border map; stage 2 walks no tile and
    applies the border map from the (all-reduced) ``xbuf``."""
    S, layouts = low.S, low.layouts
    pscal, gscal, bcoef, pt_scal = scal
    T = tables
    NS, NB, NVMAX = S.NS, S.NB, S.NVMAX
    x = np.asarray(x, dtype=float)
    lam = np.zeros(S.num_c) if lam is None else np.asarray(lam, dtype=float)
    out = dict(c=np.full(S.num_c, np.nan), dy=np.full(S.num_dy, np.nan),
               jac=np.full(S.nnz_g, np.nan), hess=np.full(S.nnz_h, np.nan),
               f=np.nan, grad=np.full(S.num_x, np.nan))
    bv = np.zeros(S.bv_size)
    nred_max = max([l.nred for l in layouts] + [1])
    partials = np.zeros((S.num_tiles, nred_max))
    sB = T["btab"]
    sec_node_all = T["sec_node"]
    tB, tE = (0, S.num_tiles) if tile_range is None else tile_range
    for tile in (range(tB, tE) if stage != 2 else ()):
        td = T["tile_desc"].reshape(-1, 8)[tile]
        q = int(td[0])
        lay, pd = layouts[q], layouts[q].pd
        fn = _node_fn(pd, NS)
        pb = T["pbase"][lay.pbase_off:lay.pbase_off + lay.pbase_size]
        ps = pscal[lay.pscal_off:lay.pscal_off + lay.pscal_size]
        o, ob = lay.ps, lay.pb
        NY, NV, NP, NQ, NF = pd.NY, pd.NV, pd.NP, pd.NQ, pd.NF
        N, K, xo, co = (int(pb[ob[k]]) for k in ("N", "K", "XOFF", "COFF"))
        sec_off = int(pb[ob["SECOFF"]])
        sec_node = sec_node_all[sec_off + q:sec_off + q + K + 1]
        k0, k1 = int(td[1]), int(td[2])
        nsec = k1 - k0
        node0 = int(td[3])
        nn = int(td[4])
        assert node0 == int(sec_node[k0]) and nn == int(sec_node[k1]) - node0 + 1
        run0, nruns = int(td[5]), int(td[6] - td[5])
        last_tile, has_prev = (k1 == K), (k0 > 0)
        sHk = [T["sec_h"][sec_off + k] if k >= 0 else 0.0 for k in range(k0 - 1, k1)]
        sOrd = [int(T["sec_order"][sec_off + k]) if k >= 0 else 0 for k in range(k0 - 1, k1)]
        assert int(td[7]) == (sOrd[0] - 1 if has_prev else 0)
        sNode = [int(sec_node[k]) - node0 if k >= 0 else 0 for k in range(k0 - 1, k1)] + [nn - 1]
        sNodeSec = [0] * nn
        for s in range(nsec):
            b, n = sNode[s + 1], sOrd[s + 1]
            for m in range(n - 1):
                sNodeSec[b + m] = s
            if s == nsec - 1:
                sNodeSec[b + n - 1] = s
        tvo = T["type_var_off"].reshape(-1, NVMAX + 1)
        prev_rows = sOrd[0] - 1 if has_prev else 0
        t0 = ps[o["TINFO"]] * (x[int(pb[ob["T0X"]])] if lay.has_t0 else 0.0) + ps[o["TINFO"] + 1]
        tF = ps[o["TINFO"] + 2] * (x[int(pb[ob["TFX"]])] if lay.has_tF else 0.0) + ps[o["TINFO"] + 3]
        hp = 0.5 * (tF - t0)
        sv = [gscal[j] * x[S.num_x - NS + j] + gscal[NS + j] for j in range(NS)]
        red = np.zeros(max(lay.nred, 1))
        nds = sum(1 for e, _ in pd.d1s if pd.fam[e] == "d")
        sF = np.zeros((NY, nn))
        sD = np.zeros((len(pd.d1v), nn))
        sDS = np.zeros((nds, nn))
        fam_code = {"d": 0, "p": 1, "i": 2}
        for ml in range(nn):
            owned = (ml < nn - 1) or last_tile
            s = sNodeSec[ml]
            mloc = ml - sNode[s + 1]
            n_k, h_k = sOrd[s + 1], sHk[s + 1]
            m = node0 + ml
            swp = (mloc == 0) and (s > 0 or has_prev)
            n_pr, h_pr = (sOrd[s], sHk[s]) if swp else (0, 0.0)
            v = [ps[o["VV"] + a] * x[xo + a * N + m] + ps[o["RV"] + a] for a in range(NV)] + sv
            wq = 0.0
            if NQ > 0:
                if swp:
                    wq = sB[T["order_w_off"][n_pr] + n_pr - 1] * h_pr
                wq = wq + sB[T["order_w_off"][n_k] + mloc] * h_k
            muh, mut = [0.0] * NF, [0.0] * NF
            if flags & F_H:
                for i in range(NY):
                    acc = 0.0
                    base = co + i * (N - 1) + node0
                    if swp:
                        Apr = T["order_a_off"][n_pr]
                        rb = sNode[s]
                        for l in range(n_pr - 1):
                            acc += lam[base + rb + l] * (sB[Apr + l * n_pr + n_pr - 1] * h_pr)
                    if owned:
                        Ak = T["order_a_off"][n_k]
                        rb = sNode[s + 1]
                        for l in range(n_k - 1):
                            acc += lam[base + rb + l] * (sB[Ak + l * n_k + mloc] * h_k)
                    mu = ps[o["WFN"] + i] * acc
                    mut[i], muh[i] = mu, hp * mu
                for j in range(NP):
                    mu = ps[o["WFN"] + NY + j] * lam[co + NY * (N - 1) + j * N + m] if owned else 0.0
                    muh[NY + j] = mu
                for i in range(NQ):
                    mu = -ps[o["WFN"] + NY + NP + i] * lam[co + NY * (N - 1) + NP * N + i] * wq
                    mut[NY + NP + i], muh[NY + NP + i] = mu, hp * mu
            R = fn(v, muh, mut)
            for i in range(NY):
                sF[i, ml] = R["F"][i]
            if (flags & F_DY) and owned:
                for i in range(NY):
                    out["dy"][int(pb[ob["DYOFF"]]) + i * N + m] = R["F"][i]
            if (flags & F_C) and owned:
                for j in range(NP):
                    out["c"][co + NY * (N - 1) + j * N + m] = ps[o["WFN"] + NY + j] * R["F"][NY + j]
            if owned:
                for i in range(NQ):
                    red[lay.red_g + i] += wq * R["F"][NY + NP + i]
            if flags & F_G:
                for k, (e, a) in enumerate(pd.d1v):
                    f = fam_code[pd.fam[e]]
                    fac = hp if f == 0 else (1.0 if f == 1 else -hp * wq)
                    sD[k, ml] = ps[o["D1V"] + k] * fac * R["D1V"][k]
                kd = kr = 0
                for k, (e, j) in enumerate(pd.d1s):
                    f = fam_code[pd.fam[e]]
                    if f == 0:
                        sDS[kd, ml] = ps[o["D1S"] + k] * hp * R["D1S"][k]
                        kd += 1
                    elif f == 1:
                        if owned:
                            out["jac"][int(pb[ob["GSCOL"] + k]) + m] = ps[o["D1S"] + k] * R["D1S"][k]
                    else:
                        if owned:
                            red[lay.red_gs + kr] += wq * R["D1S"][k]
                        kr += 1
            if (flags & F_H) and owned:
                if not (m == 0 or m == N - 1):
                    pairs_by_b = {}
                    for k, (a, b) in enumerate(pd.h2vv):
                        pairs_by_b.setdefault(b, []).append((a, k))
                    for b, lst in pairs_by_b.items():
                        lst = sorted(lst)
                        for pos, (a, k) in enumerate(lst):
                            out["hess"][int(pb[ob["HREG"] + b]) + (m - 1) * len(lst) + pos] = \
                                ps[o["H2VV"] + k] * R["H2VV"][k]
                    for k in range(len(pd.h2vs)):
                        out["hess"][int(pb[ob["HS"] + k]) + (m - 1)] = ps[o["H2VS"] + k] * R["H2VS"][k]
                    for k in range(len(pd.htv)):
                        if lay.has_t0:
                            out["hess"][int(pb[ob["HT0"] + k]) + (m - 1)] = ps[o["HT0"] + k] * R["HTV"][k]
                        if lay.has_tF:
                            out["hess"][int(pb[ob["HTF"] + k]) + (m - 1)] = ps[o["HTF"] + k] * R["HTV"][k]
                else:
                    irr = int(pb[ob["IRR0"] if m == 0 else ob["IRR1"]])
                    vals = R["H2VV"] + R["H2VS"] + R["HTV"]
                    bv[irr:irr + len(vals)] = vals
                for k in range(len(pd.hts)):
                    red[lay.red_hts + k] += R["HTS"][k]
                for k in range(len(pd.h2ss)):
                    red[lay.red_hss + k] += R["H2SS"][k]
            if (flags & F_GRAD) and owned:
                for a in range(NV):
                    out["grad"][xo + a * N + m] = 0.0
        if (flags & F_GRAD) and tile == 0:
            out["grad"][T["pt_x"]] = 0.0
        # rows
        for r in range(nn - 1):
            s = sNodeSec[r]
            b = sNode[s + 1]
            l = r - b
            n_k, h_k = sOrd[s + 1], sHk[s + 1]
            A0 = T["order_a_off"][n_k] + l * n_k
            for i in range(NY):
                acc = sum((sB[A0 + mm] * h_k) * sF[i, b + mm] for mm in range(n_k))
                if flags & F_C:
                    ya = ps[o["VV"] + i] * x[xo + i * N + node0 + b] + ps[o["RV"] + i]
                    yb = ps[o["VV"] + i] * x[xo + i * N + node0 + r + 1] + ps[o["RV"] + i]
                    out["c"][co + i * (N - 1) + node0 + r] = ps[o["WFN"] + i] * ((ya - yb) + hp * acc)
                if (flags & F_G) and pd.fn_nonzero[i]:
                    if lay.has_t0:
                        out["jac"][int(pb[ob["GT0"] + i]) + node0 + r] = ps[o["GT0"] + i] * acc
                    if lay.has_tF:
                        out["jac"][int(pb[ob["GTF"] + i]) + node0 + r] = ps[o["GTF"] + i] * acc
            if flags & F_G:
                kd = 0
                for k, (e, j) in enumerate(pd.d1s):
                    if pd.fam[e] == "d":
                        acc = sum((sB[A0 + mm] * h_k) * sDS[kd, b + mm] for mm in range(n_k))
                        out["jac"][int(pb[ob["GSCOL"] + k]) + node0 + r] = acc
                        kd += 1
        # G scatter (runs of same-type sections, flattened 64-bit recipes)
        if flags & F_G:
            cst = ps[o["GCST"]:o["GCST"] + 1 + 2 * NY]
            rgb = T["run_gbase"].reshape(-1, NVMAX)
            for r in range(nruns):
                s_lo, s_hi = int(T["run_slo"][run0 + r]), int(T["run_shi"][run0 + r])
                tv = tvo[int(T["run_type"][run0 + r])]
                rec0, Ptot = int(tv[0]), int(tv[NV] - tv[0])
                for u in range(Ptot):
                    w = int(T["recipes"][rec0 + u])
                    lo = w & 0xffffffff
                    if lo >> st.RC_SKIP_BIT:
                        continue
                    a = (w >> 32) & 0xff
                    local = (w >> st.RC_LOCAL_SHIFT) & ((1 << st.RC_LOCAL_BITS) - 1)
                    Pa = int(tv[a + 1] - tv[a])
                    e = lo & ((1 << st.RC_E_BITS) - 1)
                    bi = (lo >> st.RC_B_SHIFT) & ((1 << st.RC_B_BITS) - 1)
                    mloc = (w >> st.RC_M_SHIFT) & ((1 << st.RC_M_BITS) - 1)
                    ci = (lo >> st.RC_C_SHIFT) & ((1 << st.RC_C_BITS) - 1)
                    prev = (lo >> st.RC_PREV_BIT) & 1
                    plain = (lo >> st.RC_PLAIN_BIT) & 1
                    for sc in range(s_lo, s_hi):
                        ml = sNode[sc + 1] + mloc
                        d = sD[e - 1, ml] if e else 0.0
                        coef = 1.0 if plain else sB[bi] * (sHk[sc] if prev else sHk[sc + 1])
                        out["jac"][int(rgb[run0 + r, a]) + local + (sc - s_lo) * Pa] = coef * d + cst[ci]
        partials[tile, :len(red)] = red
    # ---- border ----
    bv[0] = 1.0
    for q, lay in enumerate(layouts):
        pb = T["pbase"][lay.pbase_off:lay.pbase_off + lay.pbase_size]
        t_lo, t_hi = int(pb[lay.pb["TILE0"]]), int(pb[lay.pb["TILE1"]])
        t_lo, t_hi = max(t_lo, tB), min(t_hi, tE)
        for k in range(lay.nred):
            bv[S.ph[q].red_off + k] = partials[t_lo:t_hi, k].sum() if t_hi > t_lo else 0.0
    if stage == 1:
        out["xbuf"] = np.concatenate([bv[1:S.bv_ptval], bv[S.bv_irr0:]])
        return out
    if stage == 2:
        nr = S.bv_ptval - 1
        bv[1:S.bv_ptval] = xbuf[:nr]
        bv[S.bv_irr0:] = xbuf[nr:nr + S.bv_size - S.bv_irr0]
    npt = len(low.ptd.pts)
    pt = [pt_scal[a] * x[T["pt_x"][a]] + pt_scal[npt + a] for a in range(npt)]
    bv[S.bv_ptval:S.bv_ptval + npt] = pt
    mult = [sigma * gscal[2 * NS]] + [
        (lam[S.num_c - NB + k] * gscal[2 * NS + 1 + k]) if (flags & F_H) else 0.0
        for k in range(NB)]
    PV, PD1, PD2 = _point_fn(low.ptd, NB)(pt, mult)
    bv[S.bv_ptfn:S.bv_ptfn + len(PV)] = PV
    bv[S.bv_ptd1:S.bv_ptd1 + len(PD1)] = PD1
    bv[S.bv_ptd2:S.bv_ptd2 + len(PD2)] = PD2
    RS = [1.0]
    for q, lay in enumerate(layouts):
        pb = T["pbase"][lay.pbase_off:lay.pbase_off + lay.pbase_size]
        ps = pscal[lay.pscal_off:lay.pscal_off + lay.pscal_size]
        i0, iF = int(pb[lay.pb["T0X"]]), int(pb[lay.pb["TFX"]])
        ti = lay.ps["TINFO"]
        t0 = ps[ti] * (x[i0] if i0 >= 0 else 0.0) + ps[ti + 1]
        tF = ps[ti + 2] * (x[iF] if iF >= 0 else 0.0) + ps[ti + 3]
        RS.append(0.5 * (tF - t0))
    want = {0: F_C, 1: F_G, 2: F_H, 3: F_J, 4: F_GRAD}
    key = {0: "c", 1: "jac", 2: "hess", 4: "grad"}
    for e in range(len(T["border_grp"])):
        grp = int(T["border_grp"][e])
        if not (flags & want[grp]):
            continue
        acc = 0.0
        for k in range(T["border_ptr"][e], T["border_ptr"][e + 1]):
            acc += bcoef[k] * bv[T["border_bv"][k]] * RS[T["border_rs"][k]]
        if grp == 3:
            out["f"] = acc
        else:
            out[key[grp]][T["border_slot"][e]] = acc
    return out

"""CUDA code generation for the per-node expression bodies.

Only the *expression bodies* are generated (task brief / BASELINE north star):
for each phase one ``__device__`` function evaluating, at one collocation node,
the state equations / path constraints / integrands, their structural first
derivatives and the multiplier-contracted second derivatives, from a common
sub-expression-eliminated straight-line program (``sympy.cse``) -- the same
idea as the reference's dead "tiered SSA" generator (``pycollo/numbafy.py:
231-245``, ``expression_graph.py``), but derivative rules come from sympy, not
from the reference's broken table.  Plus one point function
for the objective and endpoint constraints.  The generated header also carries
the constexpr dimension / offset tables the hand-written skeleton
(``csrc/pcx_kernels.cuh``) is specialised on.  Kernel code is NOT generated.

Mesh-dependent quantities never appear here, so a mesh refinement does not
trigger a recompile.
"""
from __future__ import annotations

import hashlib
import re

import sympy as sym
from sympy.printing.c import C99CodePrinter


class _CudaPrinter(C99CodePrinter):
    def _print_Float(self, expr):
        return repr(float(expr))

    def _print_Integer(self, expr):
        return f"{int(expr)}.0"

    def _print_Rational(self, expr):
        return f"({int(expr.p)}.0/{int(expr.q)}.0)"

    def _print_Pow(self, expr):
        base, exp = expr.base, expr.exp
        if exp.is_Integer and 1 <= abs(int(exp)) <= 12:
            b = self.parenthesize(base, 1000)
            prod = "*".join([b] * abs(int(exp)))
            return f"({prod})" if int(exp) > 0 else f"(1.0/({prod}))"
        if exp == sym.Rational(1, 2):
            return f"sqrt({self._print(base)})"
        if exp == sym.Rational(-1, 2):
            # IEEE division of a correctly rounded sqrt (rsqrt() is not correctly rounded)
            return f"(1.0/sqrt({self._print(base)}))"
        return f"pow({self._print(base)}, {self._print(exp)})"

    def _print_Symbol(self, expr):
        return str(expr)


_PRINTER = _CudaPrinter()
_OUTPUTS_LAST = bool(int(__import__('os').environ.get('PCX_CODEGEN_OUTPUTS_LAST', '0')))


def _ccode(e):
    return _PRINTER.doprint(sym.sympify(e))


def _cfun(name, values, default=0):
    """constexpr lookup function (arrays of static members are awkward in
    device code; a ternary chain folds away after unrolling)."""
    body = "".join(f"i=={k}?{int(v)}:" for k, v in enumerate(values))
    return (f"    static __host__ __device__ constexpr int {name}(int i) "
            f"{{ return {body}{default}; }}\n")


def _gfun(name, values, default=0):
    body = "".join(f"i=={k}?{int(v)}:" for k, v in enumerate(values))
    return (f"__host__ __device__ constexpr int {name}(int i) "
            f"{{ return {body}{default}; }}\n")


def _share_reciprocals(repl, reduced):
    """One fp64 division per distinct denominator: ``b**-n`` becomes
    ``(1/b)**n`` with ``1/b`` a shared temporary (a DDIV costs ~10x a DMUL and
    second derivatives of rational dynamics are full of ``b**-2``, ``b**-3``)."""
    recip, out = {}, []

    def fix(e):
        for pw in sorted((q for q in e.atoms(sym.Pow)
                          if q.exp.is_Integer and q.exp < 0),
                         key=sym.default_sort_key):
            b, n = pw.base, -int(pw.exp)
            if b not in recip:
                r = sym.Symbol(f"rc_{len(recip)}")
                out.append((r, sym.Pow(b, -1)))
                recip[b] = r
            e = e.xreplace({pw: recip[b] ** n})
        return e

    for s_, e in repl:
        if e.is_Pow and e.exp == -1:
            recip.setdefault(e.base, s_)
            if recip[e.base] is s_:
                out.append((s_, e))
            else:
                out.append((s_, recip[e.base]))
            continue
        out.append((s_, fix(e)))
    reduced = [fix(e) for e in reduced]
    return out, reduced


def _share_half_powers(repl, reduced):
    """``b**(p/2)`` with odd p becomes ``sqrt(b)**p`` with one shared ``sqrt(b)``
    per base: an fp64 ``pow`` is a few hundred instructions, a square root ~30, and
    gravity / drag terms are full of ``r**(-3/2)``, ``r**(-5/2)``, ``v**(3/2)``.
    Negative powers are then turned into one reciprocal by ``_share_reciprocals``."""
    roots, out = {}, []

    def fix(e):
        for pw in sorted((q for q in e.atoms(sym.Pow)
                          if q.exp.is_Rational and q.exp.q == 2 and abs(q.exp.p) >= 3),
                         key=sym.default_sort_key):
            b = pw.base
            if b not in roots:
                r = sym.Symbol(f"sq_{len(roots)}", positive=True)
                out.append((r, sym.sqrt(b)))
                roots[b] = r
            e = e.xreplace({pw: roots[b] ** int(pw.exp.p)})
        return e

    for s_, e in repl:
        out.append((s_, fix(e)))
    reduced = [fix(e) for e in reduced]
    return out, reduced


def _emit_program(inputs, outputs, indent="        "):
    """Straight-line program: ``inputs`` = [(symbol, c_expr)], ``outputs`` =
    [(c_lvalue, sympy expr)].  Shared CSE across all outputs; sin/cos of the
    same argument fused into one ``sincos``."""
    lines = []
    exprs = [sym.sympify(e) for _, e in outputs]
    # fuse sin/cos pairs on trig-free arguments
    trig_args = {}
    for e in exprs:
        for node in e.atoms(sym.sin, sym.cos):
            arg = node.args[0]
            if not arg.has(sym.sin, sym.cos):
                trig_args.setdefault(arg, set()).add(type(node))
    pre = []
    trig_sub = {}
    for k, (arg, kinds) in enumerate(sorted(trig_args.items(),
                                            key=lambda kv: sym.default_sort_key(kv[0]))):
        s_sym, c_sym = sym.Symbol(f"sn{k}"), sym.Symbol(f"cs{k}")
        if kinds == {sym.sin, sym.cos}:
            pre.append(f"{indent}double sn{k}, cs{k}; sincos({_ccode(arg)}, &sn{k}, &cs{k});")
            trig_sub[sym.sin(arg)] = s_sym
            trig_sub[sym.cos(arg)] = c_sym
        elif kinds == {sym.sin}:
            pre.append(f"{indent}const double sn{k} = sin({_ccode(arg)});")
            trig_sub[sym.sin(arg)] = s_sym
        else:
            pre.append(f"{indent}const double cs{k} = cos({_ccode(arg)});")
            trig_sub[sym.cos(arg)] = c_sym
    if trig_sub:
        exprs = [e.xreplace(trig_sub) for e in exprs]
    repl, reduced = sym.cse(exprs, symbols=sym.numbered_symbols("w_"),
                            order="none")
    repl, reduced = _share_half_powers(repl, reduced)
    repl, reduced = _share_reciprocals(repl, reduced)
    for s, cexpr in inputs:
        lines.append(f"{indent}const double {s} = {cexpr};")
    lines.extend(pre)
    # every output is emitted right after the last temporary it needs, so a
    # consumer that stores it at once (the node sink of pcx_kernels.cuh) keeps
    # no result live across the rest of the program
    where = {s_: i for i, (s_, _) in enumerate(repl)}
    after = {}
    for (lv, _), e in zip(outputs, reduced):
        pos = max([where[f] for f in e.free_symbols if f in where], default=-1)
        if _OUTPUTS_LAST:
            pos = len(repl) - 1
        stmt = lv.format(_ccode(e)) if "{}" in lv else f"{lv} = {_ccode(e)};"
        after.setdefault(pos, []).append(f"{indent}{stmt}")
    lines.extend(after.get(-1, []))
    for i, (s, e) in enumerate(repl):
        lines.append(f"{indent}const double {s} = {_ccode(e)};")
        lines.extend(after.get(i, []))
    return "\n".join(lines)


class PhaseLayout:
    """Offsets of the per-phase runtime tables (pscal: doubles, pbase: int64)."""

    def __init__(self, pd, has_t0, has_tF, pscal_off, pbase_off, kc=()):
        NV, NF, NY = pd.NV, pd.NF, pd.NY
        self.pd = pd
        # literals of a SHARED expression body that differ between the phases sharing it
        # (see share_groups): this phase's values, last block of its pscal table
        self.kc = [float(z) for z in kc]
        self.has_t0, self.has_tF = has_t0, has_tF
        o = 0
        self.ps = {}
        for name, n in (("VV", NV), ("RV", NV), ("WFN", NF), ("D1V", len(pd.d1v)),
                        ("D1S", len(pd.d1s)), ("GCST", 1 + 2 * NY),
                        ("H2VV", len(pd.h2vv)), ("H2VS", len(pd.h2vs)),
                        ("HT0", len(pd.htv)), ("HTF", len(pd.htv)),
                        ("GT0", NY), ("GTF", NY), ("TINFO", 4), ("KC", len(self.kc))):
            self.ps[name] = o
            o += n
        self.pscal_size = o
        self.pscal_off = pscal_off
        o = 0
        self.pb = {}
        for name, n in (("N", 1), ("K", 1), ("XOFF", 1), ("COFF", 1), ("DYOFF", 1),
                        ("SECOFF", 1), ("GSECOFF", 1), ("T0X", 1), ("TFX", 1),
                        ("TILE0", 1), ("TILE1", 1), ("IRR0", 1), ("IRR1", 1),
                        ("GT0", NY), ("GTF", NY), ("GSCOL", len(pd.d1s)),
                        ("HREG", NV), ("HS", len(pd.h2vs)),
                        ("HT0", len(pd.htv)), ("HTF", len(pd.htv))):
            self.pb[name] = o
            o += n
        self.pbase_size = o
        self.pbase_off = pbase_off
        # reductions (order must match structure._build_border)
        self.red_g = 0
        self.red_gs = pd.NQ
        n_gs = sum(1 for e, _ in pd.d1s if pd.fam[e] == "i")
        self.red_hts = self.red_gs + n_gs
        self.red_hss = self.red_hts + len(pd.hts)
        self.nred = self.red_hss + len(pd.h2ss)


def stage_h_rule(n_h2vv):
    """Stage the node-diagonal Hessian entries through shared memory?  Measured on
    Delta III (37 entries per node, 10^6 nodes): 0.53-0.65 ms with staging against
    0.47 ms with direct stores in every tiling tried -- the partial-sector stores are
    merged in L2 and the extra shared memory costs a resident CTA -- so the answer is
    no; ``PCX_STAGE_H=1`` keeps the experiment reproducible (correct: the GPU suite
    passes with it)."""
    import os
    return bool(int(os.environ.get("PCX_STAGE_H", "0"))) and n_h2vv > 0


_LITERAL = re.compile(r"(?<![\w.])(?:\d+\.\d*(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])")


def share_bodies_default():
    """Phases whose generated structs are textually identical up to the values of their
    floating-point literals share ONE instantiation of the tile function (Delta III: four
    phases, one 185 KB body instead of four; the literals that differ -- thrust, mass flow --
    are read from the phase's constant table).  ``PCX_SHARE_BODIES=0`` switches it off."""
    import os
    return bool(int(os.environ.get("PCX_SHARE_BODIES", "1")))


def share_groups(signatures, bodies):
    """``leader[q]`` = first phase with phase q's signature (dimension tables + body text
    with every float literal blanked); per phase the values of the literals that differ
    inside its group, in slot order; per leader the body text reading those from ``kc``."""
    P = len(bodies)
    leader = list(range(P))
    first = {}
    for q in range(P):
        leader[q] = first.setdefault(signatures[q], q)
    kc = [[] for _ in range(P)]
    shared_body = {}
    for r in sorted(set(leader)):
        group = [q for q in range(P) if leader[q] == r]
        if len(group) < 2:
            continue
        lits = [_LITERAL.findall(bodies[q]) for q in group]
        slots, slot_of = {}, {}
        for i in range(len(lits[0])):
            key = tuple(l[i] for l in lits)
            if len(set(key)) > 1:
                slot_of[i] = slots.setdefault(key, len(slots))
        for g, q in enumerate(group):
            kc[q] = [float(key[g]) for key in slots]          # dicts keep insertion order
        counter = iter(range(len(lits[0])))

        def put(m):
            i = next(counter)
            return f"kc[{slot_of[i]}]" if i in slot_of else m.group(0)
        shared_body[r] = _LITERAL.sub(put, bodies[r])
    return leader, kc, shared_body


def generate(ir, phase_derivs, point_derivs, structure, share=None):
    """Return (header_source, layouts)."""
    P = len(ir.phases)
    NS, NB = ir.n_s, ir.n_b
    npt = len(point_derivs.pts)
    share = share_bodies_default() if share is None else bool(share)
    out = ["// generated by pycollo_b200.codegen -- do not edit\n",
           "#pragma once\n",
           f"#define PCX_NUM_PHASES {P}\n#define PCX_NS {NS}\n#define PCX_NB {NB}\n",
           f"#define PCX_NPOINT {npt}\n",
           f"#define PCX_NY_MAX {max([pd.NY for pd in phase_derivs] + [1])}\n",
           f"#define PCX_NV_MAX {max([pd.NV for pd in phase_derivs] + [1])}\n",
           f"#define PCX_GS_VS 0\n#define PCX_GS_RS {NS}\n#define PCX_GS_W {2 * NS}\n"
           f"#define PCX_GS_WB {2 * NS + 1}\n",
           f"#define PCX_BV_PTVAL {structure.bv_ptval}\n#define PCX_BV_PTFN {structure.bv_ptfn}\n"
           f"#define PCX_BV_PTD1 {structure.bv_ptd1}\n#define PCX_BV_PTD2 {structure.bv_ptd2}\n"
           f"#define PCX_BV_IRR {structure.bv_irr0}\n",
           "#define PCX_FOREACH_PHASE(X) " + " ".join(f"X({q})" for q in range(P)) + "\n",
           "template <int P> struct PcxPhase;\n"]
    # the expression bodies (the expensive part: one sympy.cse program per phase), then
    # the phases that may share one
    bodies = [_cached(pd, ("body", NS), lambda pd=pd: _phase_body(pd, NS)) for pd in phase_derivs]
    leader, kcs, shared_body = list(range(P)), [[] for _ in range(P)], {}
    if share and P > 1:
        sig = [(_phase_dims(pd, PhaseLayout(pd, ph.t_needed[0], ph.t_needed[1], 0, 0)),
                _LITERAL.sub("#", bodies[q]))
               for q, (ph, pd) in enumerate(zip(ir.phases, phase_derivs))]
        leader, kcs, shared_body = share_groups(sig, bodies)
    layouts = []
    ps_off = pb_off = 0
    for q, (ph, pd) in enumerate(zip(ir.phases, phase_derivs)):
        lay = PhaseLayout(pd, ph.t_needed[0], ph.t_needed[1], ps_off, pb_off, kcs[q])
        lay.leader = leader[q]
        layouts.append(lay)
        ps_off += lay.pscal_size
        pb_off += lay.pbase_size
        out.append(_phase_struct(q, pd, lay, leader[q], q in shared_body,
                                 shared_body.get(q, bodies[q])))
    out.insert(2, f"#define PCX_PSCAL_TOTAL {ps_off}\n#define PCX_PBASE_TOTAL {pb_off}\n"
                  f"#define PCX_GSCAL_TOTAL {2 * NS + 1 + NB}\n")
    out.append(_gfun("PCX_PHASE_NRED", [l.nred for l in layouts]))
    out.append(_gfun("PCX_PHASE_REDOFF", [t.red_off for t in structure.ph]))
    out.append(_gfun("PCX_PHASE_PBASE", [l.pbase_off for l in layouts]))
    out.append(_gfun("PCX_PHASE_PSCAL", [l.pscal_off for l in layouts]))
    out.append(_gfun("PCX_PHASE_TINFO", [l.ps["TINFO"] for l in layouts]))
    out.append(_gfun("PCX_PHASE_LEADER", leader))
    pb0 = layouts[0].pb
    out.append(f"#define PCX_PB_T0X {pb0['T0X']}\n#define PCX_PB_TFX {pb0['TFX']}\n"
               f"#define PCX_PB_TILE0 {pb0['TILE0']}\n#define PCX_PB_TILE1 {pb0['TILE1']}\n")
    out.append(_cached(point_derivs, ("point", NB), lambda: _point_function(point_derivs, NB)))
    src = "".join(out)
    return src, layouts


def _cached(obj, key, make):
    """Generated text kept on the (mesh-independent, cached: derivs.analyse_phase) analysis
    object it was generated from, so that a new mesh of the same problem re-uses it."""
    store = obj.__dict__.setdefault("_generated", {})
    if key not in store:
        store[key] = make()
    return store[key]


def _phase_dims(pd, lay):
    """The constexpr members of ``PcxPhase<q>`` that do not depend on where the phase
    sits in the problem (dimensions, table offsets inside the phase's own blocks,
    pattern lookups): phases may share a body only if these agree."""
    NV, NF, NY = pd.NV, pd.NF, pd.NY
    fam_code = {"d": 0, "p": 1, "i": 2}
    pairs_by_b = {}
    for k, (a, b) in enumerate(pd.h2vv):
        pairs_by_b.setdefault(b, []).append((a, k))
    pos = [0] * len(pd.h2vv)
    nA = [0] * NV
    for b in range(NV):
        lst = sorted(pairs_by_b.get(b, []))
        nA[b] = len(lst)
        for i, (_, k) in enumerate(lst):
            pos[k] = i
    s = [f"    static constexpr int NY = {NY}, NU = {pd.NU}, NV = {NV}, NP = {pd.NP}, "
         f"NQ = {pd.NQ}, NF = {NF};\n",
         f"    static constexpr int ND1V = {len(pd.d1v)}, ND1S = {len(pd.d1s)}, "
         f"ND1SD = {sum(1 for e, _ in pd.d1s if pd.fam[e] == 'd')};\n",
         f"    static constexpr int NH2VV = {len(pd.h2vv)}, NH2VS = {len(pd.h2vs)}, "
         f"NH2SS = {len(pd.h2ss)}, NHTV = {len(pd.htv)}, NHTS = {len(pd.hts)};\n",
         f"    static constexpr bool HAS_T0 = {'true' if lay.has_t0 else 'false'}, "
         f"HAS_TF = {'true' if lay.has_tF else 'false'};\n",
         f"    static constexpr int NRED = {lay.nred}, RED_G = {lay.red_g}, "
         f"RED_GS = {lay.red_gs}, RED_HTS = {lay.red_hts}, RED_HSS = {lay.red_hss};\n"]
    for name, o in lay.ps.items():
        s.append(f"    static constexpr int OFF_{name} = {o};\n")
    for name, o in lay.pb.items():
        s.append(f"    static constexpr int PB_{name} = {o};\n")
    s.append(_cfun("FAM", [fam_code[f] for f in pd.fam]))
    s.append(_cfun("FN_NZ", [1 if z else 0 for z in pd.fn_nonzero]))
    s.append(_cfun("D1V_FN", [e for e, _ in pd.d1v]))
    s.append(_cfun("D1S_FN", [e for e, _ in pd.d1s]))
    s.append(_cfun("H2VV_B", [b for _, b in pd.h2vv]))
    s.append(_cfun("H2VV_POS", pos))
    s.append(_cfun("NA", nA))
    # experiment (off by default, see stage_h_rule): node-diagonal Hessian entries staged
    # in shared memory (one row of HP doubles per node, odd stride) and written one
    # variable block at a time, coalesced
    hb_off = [sum(nA[:b]) for b in range(NV)]
    stage_h = stage_h_rule(len(pd.h2vv))
    s.append(f"    static constexpr bool STAGE_H = {'true' if stage_h else 'false'};\n")
    s.append(f"    static constexpr int HP = {(len(pd.h2vv) | 1) if stage_h else 0};\n")
    # stride of a node's row of staged node-diagonal entries in the two-pass node phase
    s.append(f"    static constexpr int HPS = {len(pd.h2vv) | 1};\n")
    s.append(_cfun("HB_OFF", hb_off))
    return "".join(s)


def _phase_struct(q, pd, lay, leader, shared, body):
    """``PcxPhase<q>``.  ``shared``: this instantiation also serves other phases -- the
    skeleton then takes the table offsets of the phase at hand at run time instead of
    ``PSCAL_OFF`` / ``PBASE_OFF`` / ``INDEX``.  A phase with a ``leader`` other than
    itself forwards ``eval`` to the leader's body (the skeleton never instantiates it:
    it dispatches ``pcx_tile<PcxPhase<PCX_PHASE_LEADER(q)>>``)."""
    s = [f"template <> struct PcxPhase<{q}> {{\n",
         f"    static constexpr int INDEX = {q}, LEADER = {leader}, NKC = {len(lay.kc)};\n",
         f"    static constexpr bool SHARED = {'true' if shared else 'false'};\n",
         f"    static constexpr int PSCAL_OFF = {lay.pscal_off}, PBASE_OFF = {lay.pbase_off};\n",
         _phase_dims(pd, lay)]
    # results are handed to a sink (``o.template D1V<k>(value)`` ...) the moment
    # they exist instead of being returned in arrays: with tens of outputs per
    # node the arrays alone would not fit the register file
    s.append("    template <class Sink>\n"
             "    static __device__ __forceinline__ void eval(\n"
             "        const double* __restrict__ v, const double* __restrict__ muh,\n"
             "        const double* __restrict__ mut, const double* __restrict__ kc, Sink& o) {\n")
    if leader != q:
        s.append(f"        PcxPhase<{leader}>::eval(v, muh, mut, kc, o);")
    else:
        s.append(body)
    s.append("\n    }\n};\n")
    return "".join(s)


def _phase_body(pd, NS):
    """The straight-line program of one phase's node functions."""
    NV, NF = pd.NV, pd.NF
    vsyms = [sym.Symbol(f"v{a}") for a in range(NV + NS)]
    sub = dict(zip(pd.variables, vsyms))
    muh = [sym.Symbol(f"mh{e}") for e in range(NF)]
    mut = [sym.Symbol(f"mt{e}") for e in range(NF)]
    outputs = []
    for e, fe in enumerate(pd.fns):
        outputs.append((f"o.template F<{e}>({{}});", fe.xreplace(sub)))
    for k, de in enumerate(pd.d1v_expr):
        outputs.append((f"o.template D1V<{k}>({{}});", de.xreplace(sub)))
    for k, de in enumerate(pd.d1s_expr):
        outputs.append((f"o.template D1S<{k}>({{}});", de.xreplace(sub)))
    contr = {}
    for (e, a, b, dab) in pd.d2:
        contr[(a, b)] = contr.get((a, b), 0) + muh[e] * dab.xreplace(sub)
    for k, (a, b) in enumerate(pd.h2vv):
        outputs.append((f"o.template H2VV<{k}>({{}});", contr[(a, b)]))
    for k, (a, j) in enumerate(pd.h2vs):
        outputs.append((f"o.template H2VS<{k}>({{}});", contr[(a, NV + j)]))
    for k, (i, j) in enumerate(pd.h2ss):
        outputs.append((f"o.template H2SS<{k}>({{}});", contr[(NV + i, NV + j)]))
    d1_by_var = {}
    for (e, a), de in zip(pd.d1v, pd.d1v_expr):
        if pd.fam[e] in "di":
            d1_by_var[a] = d1_by_var.get(a, 0) + mut[e] * de.xreplace(sub)
    for (e, j), de in zip(pd.d1s, pd.d1s_expr):
        if pd.fam[e] in "di":
            d1_by_var[NV + j] = d1_by_var.get(NV + j, 0) + mut[e] * de.xreplace(sub)
    for k, a in enumerate(pd.htv):
        outputs.append((f"o.template HTV<{k}>({{}});", d1_by_var[a]))
    for k, j in enumerate(pd.hts):
        outputs.append((f"o.template HTS<{k}>({{}});", d1_by_var[NV + j]))
    inputs = [(f"v{a}", f"v[{a}]") for a in range(NV + NS)]
    inputs += [(f"mh{e}", f"muh[{e}]") for e in range(NF)]
    inputs += [(f"mt{e}", f"mut[{e}]") for e in range(NF)]
    return _emit_program(inputs, outputs)


def _point_function(ptd, NB):
    npt = len(ptd.pts)
    psyms = [sym.Symbol(f"p{a}") for a in range(npt)]
    sub = dict(zip(ptd.pts, psyms))
    mult = [sym.Symbol(f"m{e}") for e in range(1 + NB)]
    outputs = []
    for e, fe in enumerate(ptd.fns):
        outputs.append((f"PV[{e}]", fe.xreplace(sub)))
    for k, de in enumerate(ptd.d1_expr):
        outputs.append((f"PD1[{k}]", de.xreplace(sub)))
    contr = {}
    for (e, a, b, dab) in ptd.d2:
        contr[(a, b)] = contr.get((a, b), 0) + mult[e] * dab.xreplace(sub)
    for k, ab in enumerate(ptd.pairs):
        outputs.append((f"PD2[{k}]", contr[ab]))
    inputs = [(f"p{a}", f"pt[{a}]") for a in range(npt)]
    inputs += [(f"m{e}", f"mult[{e}]") for e in range(1 + NB)]
    body = _emit_program(inputs, outputs, indent="    ")
    return ("__device__ __noinline__ void pcx_point_eval(const double* pt, "
            "const double* mult, double* PV, double* PD1, double* PD2) {\n"
            + body + "\n}\n")


def source_hash(*parts):
    h = hashlib.sha256()
    for p in parts:
        h.update(p.encode() if isinstance(p, str) else p)
    return h.hexdigest()[:16]

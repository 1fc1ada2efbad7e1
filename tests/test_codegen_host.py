"""The generated expression bodies, compiled for the HOST with g++ and executed
(no GPU): every result of ``PcxPhase<P>::eval`` -- function values, structural
first derivatives, multiplier-contracted second derivatives, t-row sums --
against direct sympy evaluation of the same quantities at random points.

This checks the code generator's rewrites (common sub-expressions, fused sincos,
one reciprocal per denominator, half-integer powers as powers of a shared square
root, outputs interleaved with temporaries) independently of the CUDA skeleton.
"""
import os
import subprocess

import numpy as np
import pytest
import sympy as sym

from helpers import build_case
from examples import problems as examples

HARNESS = r"""
#include <cmath>
#include <cstdio>
#include <vector>
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#include "pcx_problem.h"
struct Rec {
    std::vector<double> f, d1v, d1s, h2vv, h2vs, h2ss, htv, hts;
    template <int I> void F(double v) { f[I] = v; }
    template <int K> void D1V(double v) { d1v[K] = v; }
    template <int K> void D1S(double v) { d1s[K] = v; }
    template <int K> void H2VV(double v) { h2vv[K] = v; }
    template <int K> void H2VS(double v) { h2vs[K] = v; }
    template <int K> void H2SS(double v) { h2ss[K] = v; }
    template <int K> void HTV(double v) { htv[K] = v; }
    template <int K> void HTS(double v) { hts[K] = v; }
};
template <class Ph> void run(FILE* in, int npts) {
    const int nin = Ph::NV + PCX_NS + 2 * Ph::NF + Ph::NKC;
    std::vector<double> buf(nin + 1);
    for (int k = 0; k < npts; ++k) {
        for (int i = 0; i < nin; ++i) if (fscanf(in, "%lf", &buf[i]) != 1) return;
        Rec r;
        r.f.assign(Ph::NF + 1, 0); r.d1v.assign(Ph::ND1V + 1, 0); r.d1s.assign(Ph::ND1S + 1, 0);
        r.h2vv.assign(Ph::NH2VV + 1, 0); r.h2vs.assign(Ph::NH2VS + 1, 0); r.h2ss.assign(Ph::NH2SS + 1, 0);
        r.htv.assign(Ph::NHTV + 1, 0); r.hts.assign(Ph::NHTS + 1, 0);
        Ph::eval(buf.data(), buf.data() + Ph::NV + PCX_NS, buf.data() + Ph::NV + PCX_NS + Ph::NF,
                 buf.data() + Ph::NV + PCX_NS + 2 * Ph::NF, r);
        auto dump = [](const std::vector<double>& a, int n) { for (int i = 0; i < n; ++i) printf("%.17g ", a[i]); };
        dump(r.f, Ph::NF); dump(r.d1v, Ph::ND1V); dump(r.d1s, Ph::ND1S); dump(r.h2vv, Ph::NH2VV);
        dump(r.h2vs, Ph::NH2VS); dump(r.h2ss, Ph::NH2SS); dump(r.htv, Ph::NHTV); dump(r.hts, Ph::NHTS);
        printf("\n");
    }
}
int main(int argc, char** argv) {
    FILE* in = fopen(argv[1], "r");
    int phase = 0, npts = 0;
    if (fscanf(in, "%d %d", &phase, &npts) != 2) return 1;
    switch (phase) {
#define CASE(P) case P: run<PcxPhase<P> >(in, npts); break;
        PCX_FOREACH_PHASE(CASE)
    }
    return 0;
}
"""


def _reference(pd, NS, v, muh, mut):
    NV = pd.NV
    subs = {k: sym.Float(float(z), 30) for k, z in zip(pd.variables, v)}
    ev = lambda e: float(sym.N(sym.sympify(e).xreplace(subs), 30))
    out = [ev(e) for e in pd.fns] + [ev(e) for e in pd.d1v_expr] + [ev(e) for e in pd.d1s_expr]
    contr = {}
    for (e, a, b, dab) in pd.d2:
        contr[(a, b)] = contr.get((a, b), 0.0) + muh[e] * ev(dab)
    out += [contr[ab] for ab in pd.h2vv]
    out += [contr[(a, NV + j)] for a, j in pd.h2vs]
    out += [contr[(NV + i, NV + j)] for i, j in pd.h2ss]
    d1 = {}
    for (e, a), de in zip(pd.d1v, pd.d1v_expr):
        if pd.fam[e] in "di":
            d1[a] = d1.get(a, 0.0) + mut[e] * ev(de)
    for (e, j), de in zip(pd.d1s, pd.d1s_expr):
        if pd.fam[e] in "di":
            d1[NV + j] = d1.get(NV + j, 0.0) + mut[e] * ev(de)
    out += [d1[a] for a in pd.htv] + [d1[NV + j] for j in pd.hts]
    return np.array(out)


@pytest.mark.parametrize("name", ["cart_pole_swing_up", "double_pendulum", "free_flying_robot",
                                  "space_shuttle_reentry", "delta_iii_launch_vehicle"])
def test_generated_bodies_match_sympy_on_the_host(name, tmp_path):
    low, _, _ = build_case(getattr(examples, name)(), "lobatto", 3, 4, oracle=False)
    (tmp_path / "pcx_problem.h").write_text(low.header)
    (tmp_path / "harness.cpp").write_text(HARNESS)
    exe = tmp_path / "harness"
    res = subprocess.run(["g++", "-O1", "-std=c++17", f"-I{tmp_path}", "-o", str(exe),
                          str(tmp_path / "harness.cpp"), "-lm"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    if name == "delta_iii_launch_vehicle":
        assert "pow(" not in low.header                      # half-integer powers -> shared sqrt
    rng = np.random.default_rng(3)
    NS = low.S.NS
    if name == "delta_iii_launch_vehicle":
        # the four phases differ in a handful of literals only: one shared body
        assert [lay.leader for lay in low.layouts] == [0, 0, 0, 0]
        assert 1 <= len(low.layouts[0].kc) <= 12
        assert low.header.count("        const double w_0 = ") == 1   # (the point function has its own)
    for ip in sorted({0, 1, len(low.pds) - 1} & set(range(len(low.pds)))):
        pd = low.pds[ip]
        npts = 3
        rows = []
        for _ in range(npts):
            lo, hi = (0.9, 1.6) if name == "delta_iii_launch_vehicle" else (0.2, 0.9)
            v = rng.uniform(lo, hi, pd.NV + NS)
            if name == "delta_iii_launch_vehicle":
                v[:3] *= 6.4e6                               # position outside the Earth
                v[3:6] *= 3e3
                v[6] *= 1e5
            # + the phase's own values of the literals a shared body reads from its table
            rows.append(np.concatenate([v, rng.standard_normal(pd.NF), rng.standard_normal(pd.NF),
                                        low.layouts[ip].kc]))
        inp = tmp_path / f"in_{ip}.txt"
        inp.write_text(f"{ip} {npts}\n" + "\n".join(" ".join(f"{z:.17g}" for z in r) for r in rows))
        out = subprocess.run([str(exe), str(inp)], capture_output=True, text=True, check=True).stdout
        got = np.array([[float(z) for z in line.split()] for line in out.strip().split("\n")])
        for r, g in zip(rows, got):
            v, muh, mut = (r[:pd.NV + NS], r[pd.NV + NS:pd.NV + NS + pd.NF],
                           r[pd.NV + NS + pd.NF:pd.NV + NS + 2 * pd.NF])
            ref = _reference(pd, NS, v, muh, mut)
            assert g.shape == ref.shape
            scale = np.maximum(np.abs(ref), 1e-3 * np.abs(ref).max() + 1e-300)
            assert np.max(np.abs(g - ref) / scale) <= 1e-11, (name, ip)


def test_share_groups_on_text():
    """``codegen.share_groups``: same signature => one leader; only the literals that
    differ become ``kc`` reads (one slot per distinct value tuple); identifiers with
    digits, template arguments and integers are not literals."""
    from pycollo_b200.codegen import share_groups, _LITERAL
    body = ("const double w_0 = {a}*v10 + 1.0;\n"
            "o.template D1V<3>(w_0*{b} - 2.5e-3*mh1);\n"
            "o.template F<0>({a}*w_0/{c});")
    bodies = [body.format(a="4854100.0", b="1723.25", c="3.0"),
              body.format(a="2968600.0", b="1044.5", c="3.0"),
              "const double w_0 = 2.0*v10;\no.template F<0>(w_0);",
              body.format(a="1083100.0", b="366.0", c="3.0")]
    sig = [("dims", _LITERAL.sub("#", b)) for b in bodies]
    leader, kc, shared = share_groups(sig, bodies)
    assert leader == [0, 0, 2, 0]
    assert kc == [[4854100.0, 1723.25], [2968600.0, 1044.5], [], [1083100.0, 366.0]]
    assert list(shared) == [0]
    assert shared[0] == ("const double w_0 = kc[0]*v10 + 1.0;\n"
                         "o.template D1V<3>(w_0*kc[1] - 2.5e-3*mh1);\n"
                         "o.template F<0>(kc[0]*w_0/3.0);")
    # different dimension tables never share, whatever the text
    leader, kc, shared = share_groups([("a", "x"), ("b", "x")], ["1.0", "2.0"])
    assert leader == [0, 1] and kc == [[], []] and not shared
    assert _LITERAL.findall("v10 w_12 F<3> 1.5 2e-3 7.0e+2 x1.5 3") == ["1.5", "2e-3", "7.0e+2"]


def test_symbolic_analysis_and_bodies_are_reused_across_meshes():
    """A mesh refinement re-lowers the same problem on another mesh: the derivative
    analysis and the generated bodies depend on the symbolic problem only, are found in
    the cache (same objects), and the header -- hence the compiled kernel -- is the same;
    another problem, or another derivative level, is not confused with it."""
    from pycollo_b200 import derivs
    low_a, _, _ = build_case(examples.space_shuttle_reentry(), "lobatto", 5, 4, oracle=False)
    low_b, _, _ = build_case(examples.space_shuttle_reentry(), "lobatto", 9, 6, oracle=False)
    assert low_a.pds[0] is low_b.pds[0] and low_a.ptd is low_b.ptd
    assert low_a.header == low_b.header
    assert low_a.S.num_x != low_b.S.num_x
    low_c, _, _ = build_case(examples.free_flying_robot(), "lobatto", 5, 4, oracle=False)
    assert low_c.pds[0] is not low_a.pds[0] and low_c.header != low_a.header
    ocp = examples.space_shuttle_reentry()
    ocp.settings.derivative_level = 1
    low_d, _, _ = build_case(ocp, "lobatto", 5, 4, oracle=False)
    assert low_d.pds[0] is not low_a.pds[0] and not low_d.pds[0].h2vv and low_a.pds[0].h2vv
    assert len(derivs._PHASE_CACHE) <= derivs._CACHE_MAX

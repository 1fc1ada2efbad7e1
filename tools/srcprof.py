"""Aggregate an `ncu --page source --csv` export by source line (instructions
executed, stall samples, dominant stall reasons) and print it next to the
kernel source.  usage: python tools/srcprof.py gpurun_out/src_TAG.csv [min_pct]"""
import csv, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
rows = list(csv.reader(open(path)))
# the export holds one table per source file: a header row starting "Line No"
src_k = open(os.path.join(ROOT, "pycollo_b200/csrc/pcx_kernels.cuh")).read().split("\n")
tables, cur, fname = [], None, None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = os.path.basename(r[1])
    if r and r[0] == "Line No": cur = {"file": fname, "hdr": r, "rows": []}; tables.append(cur); continue
    if cur is not None and r and r[0].isdigit(): cur["rows"].append(r)
tot_i = tot_s = 0
agg = collections.OrderedDict()
for t in tables:
    h = t["hdr"]
    ci, si = h.index("Instructions Executed"), h.index("# Samples")
    stall = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    for r in t["rows"]:
        try: n, s = int(r[ci] or 0), int(r[si] or 0)
        except ValueError: continue
        key = (t["file"], int(r[0]))
        a = agg.setdefault(key, [0, 0, collections.Counter()])
        a[0] += n; a[1] += s
        for i, nm in stall:
            try: a[2][nm[6:]] += int(r[i] or 0)
            except ValueError: pass
        tot_i += n; tot_s += s
print(f"total warp instructions {tot_i}  samples {tot_s}")
for (f, ln), (n, s, st) in sorted(agg.items()):
    if 100.0 * n / max(tot_i, 1) < minpct and 100.0 * s / max(tot_s, 1) < minpct: continue
    text = src_k[ln - 1].strip()[:70] if f and f.startswith("pcx_kernels") and ln <= len(src_k) else ""
    top = " ".join(f"{k}:{v}" for k, v in st.most_common(3) if v)
    print(f"{f}:{ln:4d} inst {100.0*n/max(tot_i,1):5.1f}%  samp {100.0*s/max(tot_s,1):5.1f}%  [{top}]  {text}")
if len(sys.argv) > 3:
    # phase totals: ranges "name:lo-hi,..." over pcx_kernels lines; pcx_problem.h counted as "eval"
    ph = collections.OrderedDict()
    for spec in sys.argv[3].split(","):
        nm, rg = spec.split(":"); lo, hi = map(int, rg.split("-")); ph[nm] = (lo, hi, [0, 0])
    ev = [0, 0]
    for (f, ln), (n, s, st) in agg.items():
        if f and f.startswith("pcx_problem"): ev[0] += n; ev[1] += s; continue
        for nm, (lo, hi, acc) in ph.items():
            if lo <= ln <= hi: acc[0] += n; acc[1] += s
    for nm, (lo, hi, acc) in ph.items():
        print(f"phase {nm:10s} inst {100.0*acc[0]/tot_i:5.1f}%  samp {100.0*acc[1]/tot_s:5.1f}%")
    print(f"phase {'eval(gen)':10s} inst {100.0*ev[0]/tot_i:5.1f}%  samp {100.0*ev[1]/tot_s:5.1f}%")

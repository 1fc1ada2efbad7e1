"""Turn the CSV exports of tools/prof.sh into the tracked summaries under profiles/.

  python tools/ncu_summary.py TAG ROUND
    gpurun_out/raw_TAG.csv       -> profiles/rROUND_pcx_fill_ncu_full.csv  (selected metrics of the full capture)
                                    profiles/traffic.json                 (DRAM bytes per launch, read by bench.py)
    gpurun_out/launches_TAG.csv  -> profiles/rROUND_launches.csv (copy) + rROUND_launches_summary.csv
    gpurun_out/src_TAG.csv       -> profiles/rROUND_source_hotspots.txt   (per-line instruction / stall shares)
"""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rnd = sys.argv[1], sys.argv[2]
name = sys.argv[3] if len(sys.argv) > 3 else "pcx_fill"      # e.g. "d3" for the Delta III capture
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEEP = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "derived__smsp__inst_executed_op_local_ld.sum",
        "smsp__inst_executed_op_local_st.sum", "smsp__inst_executed_op_local_ld.sum",
        "sm__inst_executed_pipe_lsu.sum",
        # the SM -> L2 store path: partial-sector stores travel as full sectors
        "l1tex__m_l1tex2xbar_write_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed",
        "l1tex__m_l1tex2xbar_write_sectors_mem_lg_op_st.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sectors_srcunit_tex_op_write.avg.pct_of_peak_sustained_elapsed"]
rows = list(csv.reader(open(os.path.join(G, f"raw_{tag}.csv"))))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units, data = rows[hdr], rows[hdr + 1], [r for r in rows[hdr + 2:] if len(r) == len(rows[hdr])]
out = [["metric", "unit"] + [f"launch{i}" for i in range(len(data))]]
val = {}
for k in KEEP:
    if k in names:
        j = names.index(k)
        out.append([k, units[j]] + [d[j] for d in data])
        val[k] = [d[j] for d in data]
stall = [(n, j) for j, n in enumerate(names) if n.startswith("smsp__average_warp_latency_issue_stalled") or
         n.startswith("smsp__average_warps_issue_stalled")]
for n, j in stall:
    try:
        if max(float(d[j].replace(",", "")) for d in data) >= 0.05:
            out.append([n, units[j]] + [d[j] for d in data])
    except ValueError:
        pass
dst = os.path.join(P, f"r{rnd}_{name}_ncu_full.csv")
with open(dst, "w") as fh:
    fh.write(f"# ncu --set full --clock-control none --import-source on -k regex:pcx_fill -s 8 -c 1 (tools/prof.sh {tag})\n")
    csv.writer(fh).writerows(out)


def tobytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


jr, jw = names.index("dram__bytes_read.sum"), names.index("dram__bytes_write.sum")
rd = sum(tobytes(d[jr], units[jr]) for d in data) / len(data)
wr = sum(tobytes(d[jw], units[jw]) for d in data) / len(data)
steady = os.path.join(G, f"steady_{tag}.csv")
if name != "pcx_fill":
    pass
elif os.path.exists(steady):
    # steady state: N consecutive ring launches measured in their natural cache state
    # (ncu --cache-control none, one pass: no replay), DRAM bytes averaged per launch
    acc = {"dram__bytes_read.sum": [], "dram__bytes_write.sum": [], "gpu__time_duration.sum": []}
    for r in csv.reader(open(steady)):
        if len(r) > 14 and r[12] in acc:
            acc[r[12]].append(float(r[14].replace(",", "")))
    n = len(acc["dram__bytes_read.sum"])
    srd, swr = sum(acc["dram__bytes_read.sum"]) / n, sum(acc["dram__bytes_write.sum"]) / n
    json.dump({"dram_bytes_per_launch": srd + swr, "dram_read": srd, "dram_write": swr,
               "launches_averaged": n,
               "kernel_ns_median_serialised": sorted(acc["gpu__time_duration.sum"])[n // 2],
               "cold_single_launch": {"dram_read": rd, "dram_write": wr},
               "source": f"gpurun_out/steady_{tag}.csv: ncu --cache-control none --clock-control none "
                         f"--metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:pcx_fill -s 12 -c {n} "
                         f"on `bench.py --steps 40`: {n} consecutive launches of the ring of 6 buffer sets "
                         f"(255 MB > L2), each in its natural cache state -- the write-back of earlier "
                         f"launches' values is counted where it happens, so the average is the steady-state "
                         f"DRAM traffic per evaluation.  cold_single_launch: the --set full capture (caches "
                         f"flushed before one replayed launch; its values stay in L2 until later launches)"},
              open(os.path.join(P, "traffic.json"), "w"), indent=1)
    shutil.copy(steady, os.path.join(P, f"r{rnd}_steady_state_dram.csv"))
else:
    json.dump({"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
           "source": f"profiles/r{rnd}_pcx_fill_ncu_full.csv (ncu --set full, one replayed launch: most of the "
                     f"~39 MB of values written stay in the 126 MB L2 inside one launch and are evicted "
                     f"later, so the in-launch DRAM writes are far below the algorithmic bytes)"},
              open(os.path.join(P, "traffic.json"), "w"), indent=1)
# launch list
src = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(P, f"r{rnd}_launches.csv"))
    lr = [r for r in csv.reader(open(src)) if len(r) > 10]
    h = lr[0]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    per = {}
    for r in lr[1:]:
        try:
            per.setdefault(r[kn], []).append(float(r[mv].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in per.values())
    with open(os.path.join(P, f"r{rnd}_launches_summary.csv"), "w") as fh:
        fh.write("# ncu launch list of `python bench.py --steps 30 --warmup 5 --no-cpu-baseline`\n"
                 "# (--metrics gpu__time_duration.sum --clock-control none): per-launch times are cold-cache and\n"
                 "# serialised by ncu -- compare SHARES of the step, not absolutes.  torch kernels in the list are the\n"
                 "# synthetic-input generation of bench.py's informational `amortised` section (outside every timed region)\n"
                 "kernel,launches,total_ns,median_ns,min_ns,max_ns,share_of_gpu_time\n")
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            v = sorted(v)
            fh.write(f"{k},{len(v)},{sum(v):.0f},{v[len(v)//2]:.0f},{v[0]:.0f},{v[-1]:.0f},{sum(v)/tot:.4f}\n")
hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "srcprof.py"),
                      os.path.join(G, f"src_{tag}.csv"), "0.8"], capture_output=True, text=True).stdout
open(os.path.join(P, f"r{rnd}_source_hotspots.txt" if name == "pcx_fill" else f"r{rnd}_{name}_source_hotspots.txt"), "w").write(
    "# per source line of pcx_kernels.cuh / the generated pcx_problem.h: share of executed warp instructions,\n"
    "# share of stall samples, dominant stall reasons (ncu --page source --print-source cuda,sass)\n" + hot)
print(open(dst).read()[:3000])

"""Regenerate the profiles/r2_final_* summaries from raw captures in gpurun_out/:
    python tools/summarize_profiles.py <prefix>      (gpurun_out/<prefix>_bench.json, _step_launches.csv, _stage.ncu-rep)
Needs `ncu` on PATH for the .ncu-rep pages (reads a report, no GPU)."""
import collections, csv, io, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
pre = sys.argv[1]

shutil.copy(os.path.join(G, pre + "_bench.json"), os.path.join(P, "r2_final_bench_line.json"))
shutil.copy(os.path.join(G, pre + "_step_launches.csv"), os.path.join(P, "r2_final_step_launches.csv"))

# ---- launch list -> kernel table + traffic
lines = [l for l in open(os.path.join(G, pre + "_step_launches.csv")) if not l.startswith("==")]
recs = collections.OrderedDict()
for row in csv.DictReader(lines):
    d = recs.setdefault(row["ID"], {"name": row["Kernel Name"]})
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    if row["Metric Name"] == "gpu__time_duration.sum":
        d["us"] = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    else:
        d[row["Metric Name"]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
agg = collections.OrderedDict()
for x in recs.values():
    n = re.sub(r"\(.*", "", x["name"]).replace("void hs::", "").replace("hs::", "").replace("<unnamed>::", "").replace("void at::native::", "torch ")[:46]
    a = agg.setdefault(n, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += x["us"]; a[2] += x.get("dram__bytes_read.sum", 0); a[3] += x.get("dram__bytes_write.sum", 0)
tot = sum(a[1] for a in agg.values())
out = ["# one W2 bench step (tools/prof_step.py, 4th step) under ncu --metrics gpu__time_duration.sum,dram__bytes_* --clock-control none",
       "# final round-2 code: %d launches, %.3f ms serialised (cold cache, every kernel alone on the GPU: compare shares), DRAM read %.1f MB, written %.1f MB"
       % (len(recs), tot / 1000, sum(a[2] for a in agg.values()) / 1e6, sum(a[3] for a in agg.values()) / 1e6),
       "# kernel, launches, total us, share, DRAM read MB, DRAM written MB"]
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%-46s %4d %9.1f %5.1f%% %9.1f %9.1f" % (n, a[0], a[1], 100 * a[1] / tot, a[2] / 1e6, a[3] / 1e6))
open(os.path.join(P, "r2_final_step_kernels.txt"), "w").write("\n".join(out) + "\n")
calc = sum(a[2] + a[3] for n, a in agg.items() if n.startswith("k_st_") or n.startswith("k_dv_"))
json.dump({"calctmat_stage_dram_bytes_per_step": calc, "whole_step_dram_bytes": sum(a[2] + a[3] for a in agg.values()),
           "note": "dram__bytes_read.sum + dram__bytes_write.sum summed over the k_st_* / k_dv_* launches of one W2 step (ncu launch list "
                   "profiles/r2_final_step_launches.csv, 4th step of tools/prof_step.py; ncu flushes the caches before every launch, so scratch "
                   "that stays in L2 in a real run is counted as DRAM traffic here); algorithmic I/O of the stage (inputs + packed T-matrices) is ~0.35 GB"},
          open(os.path.join(P, "r2_traffic.json"), "w"), indent=1)

# ---- ncu report -> kernel table, pipe table, hot lines
rep = os.path.join(G, pre + "_stage.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(P, "r2_final_stage_kernels.csv"), "w").write(
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py")], input=raw, capture_output=True, text=True).stdout)
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"]
short = ["us", "regs", "warps%", "issue%", "L1/smem data pipe%", "LSU inst%", "DMMA%", "FP64%", "ALU%", "FMA(int)%", "XU%", "bank conflicts", "DRAM rd MB", "DRAM wr MB"]
out = ["# ncu --set full --clock-control none: pipe utilisation of the stage kernels (round 2 of a W2 call, one particle group)" +
       (" and of the three k_scatter launches" if os.path.exists(os.path.join(G, "r2f_scatter.ncu-rep")) else ""), "# " + " | ".join(short)]
raws = [raw]
if os.path.exists(os.path.join(G, "r2f_scatter.ncu-rep")):
    raws.append(subprocess.run(["ncu", "-i", os.path.join(G, "r2f_scatter.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
for r_ in raws:
    rows = list(csv.reader(io.StringIO(r_)))
    hdr = rows[0]
    for r in rows[2:]:
        vals = []
        for w in want:
            v = r[hdr.index(w)] if w in hdr else ""
            try:
                v = "%.1f" % float(v.replace(",", ""))
            except ValueError:
                pass
            vals.append(v)
        out.append("%-46s " % r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")[:46] + " | ".join(vals))
open(os.path.join(P, "r2_final_pipes.txt"), "w").write("\n".join(out) + "\n")
txt = []
for skip, title in ((2, "k_st_assemble_q<false, 32> (wide modes m >= 1)"), (5, "k_st_assemble_q<true, 16> (narrow modes m = 0)"), (6, "k_st_mode_small_q")):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(skip), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    tmp = "/tmp/_src_%d.csv" % skip
    open(tmp, "w").write(src)
    txt.append("# %s, round 2 of a W2 call, ncu --set full --import-source on: top source lines by warp samples" % title)
    txt.append(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), tmp, "24"], capture_output=True, text=True).stdout)
open(os.path.join(P, "r2_final_assemble_lines.txt"), "w").write("\n".join(txt))
print(open(os.path.join(P, "r2_final_pipes.txt")).read())

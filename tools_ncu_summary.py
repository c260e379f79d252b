#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep: headline metrics + hottest source lines (needs -lineinfo)."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kfilt = ["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
raw = subprocess.run(["ncu", "-i", rep] + kfilt + ["--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "lts__t_bytes.sum",
        "l1tex__t_bytes.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_fp64.sum", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:75s} {units[i]:12s} {vals[i]}")
for i, h in enumerate(hdr):
    if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct"):
        try:
            v = float(vals[i])
        except ValueError:
            continue
        if v > 3:
            print(f"  stall {h.replace('smsp__warp_issue_stalled_','').replace('_per_warp_active.pct',''):40s} {v:.1f}")
stall = {}
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        try:
            stall[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(vals[i])
        except ValueError:
            pass
tot = sum(stall.values()) or 1.0
print("   stalls (share of warp-cycles): " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:7]))
# per-source-line view (cuda,sass): rows with a line number carry the aggregate of the line's SASS
src = subprocess.run(["ncu", "-i", rep] + kfilt + ["--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = []
fname, cols = "", None
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        cols = {x: i for i, x in enumerate(r)}
    elif cols and r[0] not in ("", "Function Name", "Kernel Name"):
        try:
            agg.append((int(r[cols["# Samples"]]), int(r[cols["Instructions Executed"]]), f"{fname}:{r[0]}", r[1].strip()[:110]))
        except Exception:
            pass
if agg:
    tot = sum(a[0] for a in agg) or 1
    toti = sum(a[1] for a in agg) or 1
    print(f"total samples {tot}, total warp-inst {toti}")
    for smp, ins, ln, txt in sorted(agg, reverse=True)[:topn]:
        print(f"{100*smp/tot:5.1f}% smp {100*ins/toti:5.1f}% inst  {ln}: {txt}")

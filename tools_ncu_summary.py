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
src = subprocess.run(["ncu", "-i", rep] + kfilt + ["--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    h = rows[0]
    def col(name):
        for i, x in enumerate(h):
            if x == name:
                return i
        return None
    ci, cs, cl = col("# Samples") or col("Warp Stall Sampling (All Samples)"), col("Source"), col("#")
    cinst = col("Instructions Executed")
    agg = []
    for r in rows[1:]:
        try:
            agg.append((int(r[ci]), int(r[cinst]) if cinst is not None and r[cinst] else 0, r[cl] if cl is not None else "", r[cs][:110]))
        except Exception:
            pass
    tot = sum(a[0] for a in agg) or 1
    toti = sum(a[1] for a in agg) or 1
    print(f"total samples {tot}, total inst {toti}")
    for smp, ins, ln, txt in sorted(agg, reverse=True)[:topn]:
        print(f"{100*smp/tot:5.1f}% smp {100*ins/toti:5.1f}% inst  L{ln}: {txt}")

#!/usr/bin/env python
"""Per-kernel headline metrics + stall mix of every launch in an .ncu-rep."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum']
ik = hdr.index("Kernel Name")
for v in rows[2:]:
    print("==", v[ik][:60])
    for i, h in enumerate(hdr):
        if h in want:
            print(f"   {h:68s} {rows[1][i]:10s} {v[i]}")
    items = []
    for i, h in enumerate(hdr):
        if h.startswith('smsp__pcsamp_warps_issue_stalled') and 'not_issued' not in h:
            try:
                items.append((float(v[i]), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError:
                pass
    tot = sum(x for x, _ in items) or 1
    print('   stalls:', ', '.join(f'{h} {100*x/tot:.0f}%' for x, h in sorted(items, reverse=True)[:7]))

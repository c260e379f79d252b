#!/usr/bin/env python
"""Top source lines of a kernel in an .ncu-rep by thread-instructions and stall samples (cuda,sass view)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
kf = (["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else [])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + kf, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; agg = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] in ("Function Name",) or hdr is None: continue
    if r[0] == "": continue  # sass rows
    try:
        i_s = hdr.index("# Samples"); i_t = hdr.index("Thread Instructions Executed"); i_w = hdr.index("Instructions Executed")
        agg.append((int(r[i_t]), int(r[i_w]), int(r[i_s]), cur, r[0], r[1].strip()[:100]))
    except Exception:
        pass
tt = sum(a[0] for a in agg) or 1; tw = sum(a[1] for a in agg) or 1; ts = sum(a[2] for a in agg) or 1
print(f"thread-inst {tt:.3e}  warp-inst {tw:.3e}  samples {ts}")
for a in sorted(agg, reverse=True)[:topn]:
    print(f"{100*a[0]/tt:5.1f}% tinst {100*a[1]/tw:5.1f}% winst {100*a[2]/ts:5.1f}% smp  {a[3]}:{a[4]}  {a[5]}")

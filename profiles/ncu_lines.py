#!/usr/bin/env python
"""Top source lines by warp-stall samples for one kernel of an ncu report (needs -lineinfo + --import-source on).
    python profiles/ncu_lines.py gpurun_out/prof.ncu-rep encoder_kernel [N]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur, agg, hdr = None, [], None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) > 8 and r[2] == "-":
        si = hdr.index("# Samples")
        ie = hdr.index("Instructions Executed")
        agg.append((int(r[si] or 0), cur, r[0], int(r[ie] or 0), r[1].strip()))
tot = sum(a[0] for a in agg) or 1
agg.sort(reverse=True)
print(f"total samples {tot}")
for s, f, ln, ie, src in agg[:n]:
    print(f"{100*s/tot:5.1f}%  {f}:{ln:>4}  inst={ie:<8} {src[:110]}")

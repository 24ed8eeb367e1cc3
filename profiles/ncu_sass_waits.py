#!/usr/bin/env python
"""Top SASS instructions by stall samples for one kernel, with the nearest mbarrier TRYWAIT (which barrier a spin
loop is waiting on: the smem offset identifies it).   python profiles/ncu_sass_waits.py rep kernel [N]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
data, hdr = [], None
for r in rows:
    if r and r[0] == "Address":
        hdr = r
    elif hdr and len(r) > 10 and r[0].startswith("0x"):
        data.append(r)
si = hdr.index("# Samples")
tot = sum(int(r[si] or 0) for r in data) or 1
idx = sorted(range(len(data)), key=lambda i: -int(data[i][si] or 0))[:n]
print("total samples", tot)
for i in idx:
    ctx = ""
    for j in list(range(i, max(0, i - 10), -1)) + list(range(i, min(len(data), i + 10))):
        if "TRYWAIT" in data[j][1]:
            ctx = data[j][1].strip(); break
    print(f"{100*int(data[i][si])/tot:5.1f}%  #{i:<6} {data[i][1].strip()[:60]:60} || {ctx[:70]}")

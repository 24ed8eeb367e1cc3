#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python profiles/summarize_ncu.py --tag r01_v11 --launches gpurun_out/launches.csv \
        --rep gpurun_out/prof.ncu-rep [--skip-launches N] [--stage-map]

* --launches: CSV written by `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file …`
  → profiles/<tag>_launches.md (per-kernel launch count, mean duration, share of our kernels' time).
* --rep: `ncu --set full` report → profiles/<tag>_full.md (per launch: duration, DRAM read/write bytes, DRAM %,
  tensor-pipe %, registers, top stall reasons) and profiles/roofline_traffic.json (bench.py reads it for
  `roofline.traffic`: DRAM bytes per launch of each stage's kernels).
Runs in the build container (needs the `ncu` CLI, no GPU).
"""
from __future__ import annotations

import argparse
import csv
import io
import json
import os
import re
import subprocess
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name: str) -> str:
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("void ", "")
    return name[:70]


def launches_md(path: str, tag: str) -> str:
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
        rows.append((short(r["Kernel Name"]), us, r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for n, us, g, b in rows:
        a = agg.setdefault(n, {"n": 0, "us": 0.0, "grid": g, "block": b})
        a["n"] += 1
        a["us"] += us
    ours = {k: v for k, v in agg.items() if not k.startswith("at::") and "profile_delay_kernel" not in k}
    tot = sum(v["us"] for v in ours.values())
    out = [f"# ncu launch list — {tag}", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: compare",
           "SHARES, not absolutes).  Share = of the time of this library's kernels in the capture.", "",
           "| kernel | launches | mean us | total us | share | grid | block |", "|---|---|---|---|---|---|---|"]
    for k, v in agg.items():
        sh = f"{100 * v['us'] / tot:.1f}%" if k in ours and tot else "-"
        out.append(f"| `{k}` | {v['n']} | {v['us'] / v['n']:.2f} | {v['us']:.1f} | {sh} | {v['grid']} | {v['block']} |")
    return "\n".join(out) + "\n"


WANT = {
    "gpu__time_duration.sum": "dur",
    "dram__bytes_read.sum": "rd",
    "dram__bytes_write.sum": "wr",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active": "tensor_inst_pct",
    "launch__registers_per_thread": "regs",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ",
    "lts__t_sector_hit_rate.pct": "l2_hit",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "launch__grid_size": "grid",
}


def num(x: str):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def to_bytes(v: float, unit: str) -> float:
    u = unit.lower()
    m = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}
    return v * m.get(u, 1)


def to_us(v: float, unit: str) -> float:
    u = unit.lower()
    m = {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
    return v * m.get(u, 1)


def full_md(rep: str, tag: str):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    hdr, units, rows = rd[0], rd[1], rd[2:]
    col = {}
    for i, h in enumerate(hdr):
        base = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1] == "TriageCompute" else h
        for w in WANT:
            if h == w or base == w:
                col.setdefault(WANT[w], i)
    tensor_cols = [i for i, h in enumerate(hdr) if "pipe_tensor" in h and "pct" in h]
    # pc-sampling stall counters (share of samples per launch)
    stall_cols = [i for i, h in enumerate(hdr) if re.fullmatch(r"smsp__pcsamp_warps_issue_stalled_[a-z_]+", h) and not h.endswith("_not_issued")]
    ki = hdr.index("Kernel Name")
    out = [f"# ncu --set full — {tag}", "",
           "One forward (B=1024), `--clock-control none`, ≈40 replay passes per launch (durations are under the",
           "profiler: use for SHARES and for the per-launch DRAM traffic; bench.py times the real thing).", "",
           "| # | kernel | grid | us | DRAM read MB | DRAM write MB | DRAM % | tensor pipe % | regs | top stalls |",
           "|---|---|---|---|---|---|---|---|---|---|"]
    recs = []
    for n, r in enumerate(rows):
        def g(key, conv=None):
            i = col.get(key)
            if i is None or num(r[i]) is None:
                return None
            v = num(r[i])
            return conv(v, units[i]) if conv else v
        tp = max([num(r[i]) for i in tensor_cols if num(r[i]) is not None] or [0.0])
        sv = [(num(r[i]) or 0.0, hdr[i].replace("smsp__pcsamp_warps_issue_stalled_", "")) for i in stall_cols]
        ssum = sum(v for v, _ in sv) or 1.0
        st = sorted(((100 * v / ssum, nm) for v, nm in sv), reverse=True)[:3]
        rec = {"kernel": short(r[ki]), "us": g("dur", to_us), "rd": g("rd", to_bytes), "wr": g("wr", to_bytes),
               "dram_pct": g("dram_pct"), "tensor_pct": tp, "regs": g("regs"), "grid": g("grid"),
               "stalls": ", ".join(f"{nm} {v:.0f}%" for v, nm in st)}
        recs.append(rec)
        out.append(f"| {n} | `{rec['kernel']}` | {int(rec['grid'] or 0)} | {rec['us']:.1f} | {(rec['rd'] or 0) / 1e6:.1f} | "
                   f"{(rec['wr'] or 0) / 1e6:.1f} | {rec['dram_pct'] or 0:.1f} | {tp:.1f} | {int(rec['regs'] or 0)} | {rec['stalls']} |")
    tot = sum(r["us"] for r in recs)
    out += ["", f"Sum of durations: {tot:.1f} us; DRAM read {sum(r['rd'] or 0 for r in recs) / 1e6:.1f} MB, "
                f"write {sum(r['wr'] or 0 for r in recs) / 1e6:.1f} MB per forward."]
    return "\n".join(out) + "\n", recs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", required=True)
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    ap.add_argument("--stages", help="JSON: stage name -> list of 0-based launch indices in the --rep capture")
    a = ap.parse_args()
    pdir = os.path.join(ROOT, "profiles")
    if a.launches:
        with open(os.path.join(pdir, f"{a.tag}_launches.md"), "w") as f:
            f.write(launches_md(a.launches, a.tag))
    if a.rep:
        md, recs = full_md(a.rep, a.tag)
        with open(os.path.join(pdir, f"{a.tag}_full.md"), "w") as f:
            f.write(md)
        if a.stages:
            smap = json.loads(a.stages)
            traffic = {}
            for stage, idx in smap.items():
                sel = [recs[i] for i in idx if i < len(recs)]
                if sel:   # DRAM bytes per launch of the stage (a stage can hold several launches: sum / count of stage runs)
                    traffic[stage] = round(sum((r["rd"] or 0) + (r["wr"] or 0) for r in sel))
            traffic["_source"] = f"profiles/{a.tag}_full.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"
            with open(os.path.join(pdir, "roofline_traffic.json"), "w") as f:
                json.dump(traffic, f, indent=1)


if __name__ == "__main__":
    main()

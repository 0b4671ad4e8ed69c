#!/usr/bin/env python
"""Turns `ncu --set full` captures into profiles/traffic.json and a markdown table (no GPU needed).

    python tools/ncu_to_traffic.py --workload C3 gpurun_out/prof_c3.ncu-rep [--workload C2 other.ncu-rep ...] --md profiles/r02_ncu_c3.md

For every kernel of a report: launches captured, mean duration, mean dram__bytes_read.sum + dram__bytes_write.sum per
launch (the `roofline.traffic` of bench.py), dram throughput %, issue-slot %, registers, warp instructions, L2 hit rate.
bench.py reads profiles/traffic.json and reports `traffic_source` = the `_source` string written here (engine commit +
report names), so the figure in the JSON line is always traceable to a committed capture and goes stale visibly.
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

METRICS = {
    "gpu__time_duration.sum": "ns",
    "dram__bytes_read.sum": "rd",
    "dram__bytes_write.sum": "wr",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_pct",
    "launch__registers_per_thread": "regs",
    "smsp__inst_executed.sum": "winst",
    "lts__t_sector_hit_rate.pct": "l2hit",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9,
              "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}


def short_name(full):
    m = re.search(r"(k_[A-Za-z0-9_]+)", full)
    return m.group(1) if m else full.split("(")[0][:40]


def read_report(path):
    """`path`: a .ncu-rep, or the text of `ncu -i <rep> --page raw --csv` saved on the GPU box (large reports stay there)"""
    if path.endswith(".csv"):
        text = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
        if out.returncode != 0:
            raise SystemExit(f"ncu -i {path} failed: {out.stderr[:400]}")
        text = out.stdout
    rows = list(csv.reader(io.StringIO(text)))
    hdr = None
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr = i
            break
    if hdr is None:
        raise SystemExit(f"{path}: no kernel table")
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: k for k, n in enumerate(names)}
    kn = col["Kernel Name"]
    per = {}
    for r in rows[hdr + 2:]:
        if len(r) <= kn:
            continue
        k = short_name(r[kn])
        d = per.setdefault(k, {v: [] for v in METRICS.values()})
        for m, key in METRICS.items():
            if m in col and r[col[m]] not in ("", "n/a"):
                try:
                    val = float(r[col[m]].replace(",", ""))
                except ValueError:
                    continue
                val *= UNIT_SCALE.get(units[col[m]], 1.0) if key in ("ns", "rd", "wr") else 1.0
                d[key].append(val)
    return per


def mean(v):
    return sum(v) / len(v) if v else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", action="append", nargs=2, metavar=("KEY", "REPORT"), required=True)
    ap.add_argument("--md", default=None)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "traffic.json"))
    args = ap.parse_args()
    try:
        traffic = json.load(open(args.out))
    except Exception:
        traffic = {}
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    md = []
    sources = []
    for key, rep in args.workload:
        per = read_report(rep)
        sources.append(f"{key}:{os.path.basename(rep)}")
        traffic.setdefault(key, {}).update({k: (mean(d["rd"]) or 0) + (mean(d["wr"]) or 0) for k, d in per.items() if d["rd"] or d["wr"]})
        md.append(f"### {key} -- `{os.path.basename(rep)}` (ncu --set full, per launch, means over the captured launches)\n")
        md.append("| kernel | launches | us | dram read MB | dram write MB | dram % | issue slots % | SM % | warps active % | regs | warp instr | L2 hit % |")
        md.append("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
        for k, d in sorted(per.items(), key=lambda kv: -(mean(kv[1]["ns"]) or 0) * len(kv[1]["ns"])):
            f = lambda x, s=1.0, p=1: "-" if mean(d[x]) is None else f"{mean(d[x]) / s:.{p}f}"
            md.append(f"| `{k}` | {len(d['ns'])} | {f('ns', 1e3)} | {f('rd', 1e6)} | {f('wr', 1e6)} | {f('dram_pct')} | {f('issue_pct')} | {f('sm_pct')} | {f('warps_pct')} | {f('regs', 1, 0)} | {f('winst', 1, 0)} | {f('l2hit')} |")
        md.append("")
    traffic["_source"] = f"ncu --set full captures {', '.join(sources)} (tools/ncu_to_traffic.py at {head}); dram__bytes_read.sum + dram__bytes_write.sum per launch"
    with open(args.out, "w") as fh:
        json.dump(traffic, fh, indent=1, sort_keys=True)
    text = "\n".join(md)
    if args.md:
        with open(args.md, "a") as fh:
            fh.write(text + "\n")
    print(text)


if __name__ == "__main__":
    sys.exit(main())

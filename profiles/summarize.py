#!/usr/bin/env python
"""Turns an ncu launch list (--metrics gpu__time_duration.sum --csv) and a raw page export
(ncu -i X.ncu-rep --page raw --csv) into the markdown tables kept under profiles/.
usage: summarize.py launches.csv raw.csv > summary.md"""
import csv
import sys
from collections import OrderedDict

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        name, val = r[4], float(r[-1])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += val
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | avg us | share of captured time |\n|---|---|---|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {t / c / 1e3:.1f} | {100 * t / tot:.1f}% |")


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    seen = OrderedDict()
    for r in data:
        seen.setdefault(r[ki], r)     # first capture of each kernel
    names = list(seen)
    print("\n| metric | " + " | ".join(f"`{n}`" for n in names) + " |\n|---|" + "---|" * len(names))
    for m in KEEP:
        if m in hdr:
            i = hdr.index(m)
            print(f"| {m} [{units[i]}] | " + " | ".join(seen[n][i] for n in names) + " |")


if __name__ == "__main__":
    launches(sys.argv[1])
    if len(sys.argv) > 2:
        raw(sys.argv[2])

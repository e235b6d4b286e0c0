#!/usr/bin/env python
"""Condense the ncu CSV exports of tools/gpu_run_a.sh (gpurun_out/) into the tracked summaries
under profiles/:  <tag>_launch_shares.txt, <tag>_ncu_full_summary.txt, dram_traffic.json.

usage: python tools/summarize_ncu.py r1b
"""
import collections
import csv
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

SLOT = [  # kernel name prefix -> bench.py profile slot
    ("k_dij", "dij"), ("k_alpha", "alpha"), ("k_low_order", "low_order"), ("k_pij_lij", "pij_lij"),
    ("k_high_order<Euler<3>, 0>", "high_order_next_lij"), ("k_high_order<Euler<3>, 1>", "high_order_final"),
]

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("launch__registers_per_thread", "registers"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math pipe"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not selected"),
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"rjb::", "", name)
    return re.sub(r"\(.*", "", name)


def launch_shares(tag):
    src = os.path.join(OUT, "launches.csv")
    rows = list(csv.reader(open(src)))
    start = next(n for n, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        agg.setdefault(short(r[ki]), []).append(v)
    total = sum(sum(v) for v in agg.values())
    lines = ["# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 1",
             "# (8.0M DoFs, 1 GPU; per-launch times are serialised by the profiler: the SHARES are what counts)",
             "%-62s %5s %10s %8s" % ("kernel", "n", "avg ms", "share")]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append("%-62s %5d %10.3f %7.1f%%" % (k[:62], len(v), sum(v) / len(v), 100 * sum(v) / total))
    open(os.path.join(PROF, tag + "_launch_shares.txt"), "w").write("\n".join(lines) + "\n")
    shutil.copy(src, os.path.join(PROF, tag + "_launches_bench_steps2.csv"))


def full_summary(tag):
    rows = list(csv.reader(open(os.path.join(OUT, "prof_%s_raw.csv" % tag))))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    seen, lines, traffic = set(), [], collections.defaultdict(list)
    lines.append("# ncu --set full --clock-control none, python bench.py --steps 1 --warmup 1 (8.0M DoFs, 1 GPU)")
    for r in rows[2:]:
        name = short(r[ki])
        rd = float(r[hdr.index("dram__bytes_read.sum")]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_read.sum")]]
        wr = float(r[hdr.index("dram__bytes_write.sum")]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_write.sum")]]
        for prefix, slot in SLOT:
            if name.startswith(prefix):
                traffic[slot].append(rd + wr)
        if name in seen:
            continue
        seen.add(name)
        lines.append("")
        lines.append(name)
        for m, label in METRICS:
            if m in hdr:
                lines.append("    %-28s %14s %s" % (label, r[hdr.index(m)], units[hdr.index(m)]))
        lines.append("    %-28s %14.3f GB  (%.0f B per DoF)" % ("dram read+write", (rd + wr) / 1e9, (rd + wr) / 8.0e6))
        t_ms = float(r[hdr.index("gpu__time_duration.sum")]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[hdr.index("gpu__time_duration.sum")]]
        lines.append("    %-28s %14.1f GB/s" % ("dram bandwidth", (rd + wr) / 1e9 / (t_ms * 1e-3)))
    open(os.path.join(PROF, tag + "_ncu_full_summary.txt"), "w").write("\n".join(lines) + "\n")
    json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(os.path.join(PROF, "dram_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1b"
    launch_shares(tag)
    full_summary(tag)
    print(open(os.path.join(PROF, tag + "_ncu_full_summary.txt")).read()[:1500])
    print(open(os.path.join(PROF, "dram_traffic.json")).read())

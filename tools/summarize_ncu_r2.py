#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export (one row per captured kernel) into a readable summary
and into profiles/fp64_instructions.json (FP64 warp-instructions per DoF of the two FP64-issue bound
sweeps, derived from the measured FP64 pipe utilisation) and profiles/dram_traffic.json.

usage: python tools/summarize_ncu_r2.py gpurun_out/r2d_full_raw.csv r2d 8000000
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, tag, n_dofs = sys.argv[1], sys.argv[2], float(sys.argv[3])

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("launch__registers_per_thread", "registers"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe % (of its issue peak)"),
    ("sm__inst_executed.sum.per_cycle_active", "warp inst / cycle (all SMs)"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math pipe"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not selected"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("sm__cycles_active.avg", "sm cycles active (avg)"),
]
SLOT = [("k_dij", "dij"), ("k_alpha", "alpha"), ("k_low_order", "low_order"), ("k_pij_lij", "pij_lij"),
        ("k_high_order_tile", "high_order_next_lij"), ("k_high_order_record", "high_order_final"),
        ("k_high_order<", "high_order_final"), ("k_dii_tau", "dii_tau")]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"rjb::", "", name)
    return re.sub(r"\(.*", "", name)


rows = list(csv.reader(open(src)))
hdr, units = rows[0], rows[1]
col = {h: n for n, h in enumerate(hdr)}
seen, out, traffic, fp64 = set(), [], {}, {}
for r in rows[2:]:
    name = short(r[col["Kernel Name"]])
    if name in seen:
        continue
    seen.add(name)
    out.append(name)
    vals = {}
    for metric, label in METRICS:
        if metric not in col:
            continue
        raw, unit = r[col[metric]].replace(",", ""), units[col[metric]]
        try:
            v = float(raw)
        except ValueError:
            continue
        vals[metric] = (v, unit)
        out.append("    %-34s %14.6g %s" % (label, v, unit))
    t, tu = vals["gpu__time_duration.sum"]
    t_s = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[tu]
    rd, ru = vals["dram__bytes_read.sum"]
    wr, wu = vals["dram__bytes_write.sum"]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total = rd * scale[ru] + wr * scale[wu]
    out.append("    %-34s %14.1f B per DoF   (%.0f GB/s)" % ("dram read+write", total / n_dofs, total / t_s / 1e9))
    slot = next((s for p, s in SLOT if name.startswith(p)), None)
    if slot and slot not in traffic:
        traffic[slot] = total
    # FP64 warp-instructions: pipe utilisation x issue peak (2 per cycle per SM) x active cycles x 148 SMs
    util = vals.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", (0.0, ""))[0] / 100.0
    cycles = vals.get("sm__cycles_active.avg", (0.0, ""))[0]
    n_inst = util * 2.0 * cycles * 148
    out.append("    %-34s %14.1f per DoF" % ("fp64 warp-inst (x32 lanes)", n_inst * 32 / n_dofs))
    if slot in ("dij", "low_order"):
        key = slot if slot == "dij" else name
        fp64.setdefault(slot, []).append(n_inst / n_dofs)
    out.append("")
open(os.path.join(ROOT, "profiles", "%s_ncu_full_summary.txt" % tag), "w").write(
    "# ncu --set full --clock-control none, python bench.py --steps 1 --warmup 1 (8.0M DoFs, 1 GPU), first launch of every kernel\n\n" + "\n".join(out))
json.dump({k: v for k, v in traffic.items()}, open(os.path.join(ROOT, "profiles", "dram_traffic.json"), "w"), indent=1)
json.dump({slot: {"fp64_warp_instructions_per_dof": sum(v) / len(v),
                  "source": "profiles/%s_ncu_full_summary.txt (fp64 pipe utilisation x 2 inst/cycle/SM x active cycles; "
                            "low_order: mean over stages 0, 1, 2)" % tag} for slot, v in fp64.items()},
          open(os.path.join(ROOT, "profiles", "fp64_instructions.json"), "w"), indent=1)
print("\n".join(out))

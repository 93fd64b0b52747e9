#!/usr/bin/env python
"""Condense `ncu --page raw --csv` exports into profiles/<tag>_ncu_summary.{json,txt}.
usage: summarize_ncu.py TAG NCOL name=raw.csv [name=raw.csv ...]   (name = the bench profiler's kernel name)"""
import csv
import json
import sys

tag, ncol = sys.argv[1], int(sys.argv[2])
KEEP = [
    ("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("smsp__inst_executed.sum", "warp_insts"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_active_pct"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_pipe"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_instruction"),
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}
out, lines = {}, []
for arg in sys.argv[3:]:
    name, path = arg.split("=", 1)
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    rec = {"kernel_function": d.get("Kernel Name", ("?", ""))[0], "ncol": ncol, "launches": 1}
    for key, short in KEEP:
        if key in d:
            v, u = d[key]
            rec[short] = float(v.replace(",", "")) * UNIT.get(u, 1.0)
    rec["dram_bytes"] = rec.get("dram_read", 0.0) + rec.get("dram_write", 0.0)
    rec["dram_bytes_per_column"] = rec["dram_bytes"] / ncol
    rec["dram_gbs"] = rec["dram_bytes"] / rec["time"] / 1e9
    out[name] = rec
    lines.append(f"== {name}  [{rec['kernel_function'][:90]}]  ({ncol} columns, one launch, ncu --set full --clock-control none)")
    for k in ("time", "regs", "warps_active_pct", "dram_read", "dram_write", "dram_gbs", "dram_bytes_per_column", "l1tex_pct",
              "l2_pct", "warp_insts", "issue_active_pct", "fp64_pipe_active_pct", "stall_long_scoreboard",
              "stall_short_scoreboard", "stall_wait", "stall_math_pipe", "stall_barrier", "stall_no_instruction"):
        if k in rec:
            lines.append(f"   {k:26s} {rec[k]:.6g}")
json.dump(out, open(f"profiles/{tag}_ncu_summary.json", "w"), indent=1)
open(f"profiles/{tag}_ncu_summary.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

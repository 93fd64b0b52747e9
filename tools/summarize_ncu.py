#!/usr/bin/env python
"""Condense `ncu --page raw --csv` exports into profiles/<tag>_ncu_summary.{json,txt}.
usage: summarize_ncu.py TAG NCOL name=raw.csv [name=raw.csv ...]   (name = the bench profiler's kernel name)
       summarize_ncu.py TAG NCOL all=raw.csv      one export holding several kernels: rows are named by kernel function
                                                   (gas_tau_g_kernel<0,..> -> gas_tau_fused[lw], <1,..> -> [sw], ...)"""
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
    # the L1/shared -> register-file return path (shared-memory loads use it too) and spills
    ("l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1_to_rf_writeback_pct"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1_data_pipe_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pipe_pct"),
    ("smsp__inst_executed_op_local_ld.sum", "local_loads"), ("smsp__inst_executed_op_local_st.sum", "local_stores"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
]


def bench_name(fn):
    if "gas_tau_g_kernel" in fn:
        return "gas_tau_fused[sw]" if "gas_tau_g_kernel<1" in fn or "gas_tau_g_kernel<(bool)1" in fn else "gas_tau_fused[lw]"
    if "planck_g_kernel" in fn:
        return "planck_fused"
    for n in ("sw_2stream_reg_kernel", "lw_noscat_reg_kernel", "lw_2stream_reg_kernel", "lw_rescl_reg_kernel"):
        if n in fn:
            return n
    return fn.split("(")[0][:40]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}
out, lines = {}, []
for arg in sys.argv[3:]:
    name0, path = arg.split("=", 1)
    rows = [r for r in csv.reader(open(path)) if r]
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        name = bench_name(d.get("Kernel Name", ("?", ""))[0]) if name0 == "all" else name0
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
                  "stall_short_scoreboard", "stall_wait", "stall_math_pipe", "stall_barrier", "stall_no_instruction",
                  "l1_to_rf_writeback_pct", "l1_data_pipe_pct", "lsu_pipe_pct", "local_loads", "local_stores", "l1_hit_pct",
                  "l2_hit_pct"):
            if k in rec:
                lines.append(f"   {k:26s} {rec[k]:.6g}")
json.dump(out, open(f"profiles/{tag}_ncu_summary.json", "w"), indent=1)
open(f"profiles/{tag}_ncu_summary.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

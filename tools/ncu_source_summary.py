#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: opcode mix (share of executed warp instructions and of
stall samples) and the most-stalled SASS lines.  Usage: ncu_source_summary.py file.csv [ntop]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 20
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
idx = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[idx["Instructions Executed"]].isdigit()]
I, S, SRC, TH = idx["Instructions Executed"], idx["# Samples"], idx["Source"], idx["Avg. Threads Executed"]
tot = sum(int(r[I]) for r in data) or 1
samp = sum(int(r[S]) for r in data) or 1
print(f"kernel: {rows[0][1] if rows[0] else '?'}")
print(f"SASS lines {len(data)}  warp instructions {tot}  stall samples {samp}")
ops, ops_s = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[SRC])
    op = m.group(2).split(".")[0] if m else "?"
    ops[op] += int(r[I])
    ops_s[op] += int(r[S])
for op, c in ops.most_common(22):
    print(f"  {op:10s} inst {c / tot * 100:5.1f}%   samples {ops_s[op] / samp * 100:5.1f}%")
print("most-stalled lines: samples, executed, avg threads, SASS")
for r in sorted(data, key=lambda r: -int(r[S]))[:ntop]:
    print(f"  {r[S]:>7s} {r[I]:>10s} {r[TH]:>5s}  {r[SRC][:100]}")

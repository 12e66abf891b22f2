#!/usr/bin/env python
"""Basic-block view of an ncu source page: consecutive SASS instructions with the same execution
count are merged into one row (first/last opcode, #instructions, executions, average active threads,
stall samples).  usage: ncu_blocks.py <report.ncu-rep> [tiles]   (tiles: divide executions by it)"""
import csv
import subprocess
import sys

path = sys.argv[1]
tiles = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
text = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(text.splitlines()))
hdr = rows[1]
c = {h: i for i, h in enumerate(hdr)}
blocks = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    ex = int(r[c["Instructions Executed"]])
    th = float(r[c["Avg. Threads Executed"]] or 0)
    smp = int(r[c["# Samples"]] or 0)
    op = r[c["Source"]].strip()
    if blocks and blocks[-1]["ex"] == ex:
        b = blocks[-1]
        b["n"] += 1
        b["smp"] += smp
        b["th"] += th
        b["last"] = op
        b["ops"].append(op.split()[0] if not op.startswith("@") else op.split()[1])
    else:
        blocks.append({"ex": ex, "n": 1, "smp": smp, "th": th, "first": op, "last": op, "ops": [op.split()[0] if not op.startswith("@") else op.split()[1]]})
total = sum(b["ex"] * b["n"] for b in blocks)
tsmp = sum(b["smp"] for b in blocks)
print("total warp-instructions %d (%.1f per tile), samples %d" % (total, total / tiles, tsmp))
print("%6s %9s %8s %6s %7s  %s" % ("#instr", "exec/tile", "instr/t", "thr", "smp%", "first .. last"))
for b in blocks:
    if b["ex"] == 0:
        continue
    w = b["ex"] * b["n"]
    if w / total < 0.002 and b["smp"] / max(tsmp, 1) < 0.002:
        continue
    print("%6d %9.3f %8.2f %6.1f %6.1f%%  %s .. %s" % (b["n"], b["ex"] / tiles, w / tiles, b["th"] / b["n"], 100.0 * b["smp"] / max(tsmp, 1),
                                                      b["first"][:48], b["last"][:40]))

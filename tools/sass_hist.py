#!/usr/bin/env python
"""Opcode histogram of one kernel from an `ncu --page source --csv` dump (SASS view): executed warp-instructions per
opcode, and per-pixel counts when --per N is given.  usage: ncu -i X.ncu-rep --page source --csv --kernel-name K | sass_hist.py [--per N]"""
import csv
import sys
from collections import Counter

per = float(sys.argv[sys.argv.index("--per") + 1]) if "--per" in sys.argv else None
rows = list(csv.reader(sys.stdin))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
si, ei, ti = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
ops, tot, tt = Counter(), 0, 0
for r in rows[h + 1:]:
    if len(r) <= ti or not r[ei].isdigit():
        continue
    parts = r[si].split()
    op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
    op = op.split(".")[0]
    n = int(r[ei])
    ops[op] += n
    tot += n
    tt += int(r[ti])
print(f"warp-instructions {tot}  thread-instructions {tt}" + (f"  per unit {tt / per:.0f}" if per else ""))
for op, n in ops.most_common(28):
    print(f"{op:12s} {n:12d} {100.0 * n / tot:5.1f} %")

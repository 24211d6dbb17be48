#!/usr/bin/env python
"""Warp-stall samples and executed instructions of one kernel per PHASE, phases being the stretches of SASS between block-wide
barriers (BAR.SYNC), from an `ncu --page source --csv` dump.  usage: ncu -i X.ncu-rep --page source --csv --kernel-name K | phase_samples.py"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
si, ni, ei = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, k) for i, k in enumerate(hdr) if k.startswith("stall_") and "Not Issued" not in k]
phases, cur = [], {"n": 0, "samples": 0, "exec": 0, "stalls": {}}
for r in rows[h + 1:]:
    if len(r) <= ei or not r[ei].isdigit():
        continue
    cur["n"] += 1
    cur["samples"] += int(r[ni])
    cur["exec"] += int(r[ei])
    for i, k in stall_cols:
        cur["stalls"][k] = cur["stalls"].get(k, 0) + int(r[i] or 0)
    if "BAR.SYNC" in r[si]:
        phases.append(cur)
        cur = {"n": 0, "samples": 0, "exec": 0, "stalls": {}}
phases.append(cur)
tot_s, tot_e = sum(p["samples"] for p in phases), sum(p["exec"] for p in phases)
print(f"total: {tot_s} samples, {tot_e} warp-instructions")
for j, p in enumerate(phases):
    top = sorted(p["stalls"].items(), key=lambda kv: -kv[1])[:4]
    print(f"phase {j}: {p['n']:5d} SASS lines  {100.0 * p['samples'] / max(tot_s, 1):5.1f} % of samples  {100.0 * p['exec'] / max(tot_e, 1):5.1f} % of instructions   "
          + "  ".join(f"{k[6:]}={100.0 * v / max(p['samples'], 1):.0f}%" for k, v in top))

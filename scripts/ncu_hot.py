#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: executed warp instructions, shared-memory wavefronts and stall samples
per contiguous SASS region (regions split where the executed count changes by more than 2x), plus opcode totals.
usage: python scripts/ncu_hot.py source.csv [pair_steps]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    ins.append((r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]] or 0), int(r[ix["L1 Wavefronts Shared"]] or 0),
                int(r[ix["# Samples"]] or 0)))
total = sum(i[1] for i in ins)
print("total warp instructions", total, "" if not steps else f"= {total / steps:.2f} per pair step")
print("total shared wavefronts", sum(i[2] for i in ins), "" if not steps else f"= {sum(i[2] for i in ins) / steps:.2f} per pair step")
# regions
regions, cur = [], None
for k, (src, ex, wf, smp) in enumerate(ins):
    if cur is None or not (0.5 * cur["ex0"] <= ex <= 2.0 * cur["ex0"]):
        cur = dict(start=k, ex0=max(ex, 1), n=0, ex=0, wf=0, smp=0, ops=collections.Counter())
        regions.append(cur)
    cur["n"] += 1; cur["ex"] += ex; cur["wf"] += wf; cur["smp"] += smp
    cur["ops"][src.split()[1].split(".")[0] if src.startswith("@") else src.split()[0].split(".")[0]] += ex
regions.sort(key=lambda r: -r["ex"])
for r in regions[:14]:
    top = ", ".join(f"{k}:{v / r['ex']:.2f}" for k, v in r["ops"].most_common(8))
    print(f"@{r['start']:5d} +{r['n']:4d} instr  exec {r['ex'] / total:6.1%}  wavefronts {r['wf']:.3e}  samples {r['smp']:7d}  "
          f"per-instr count {r['ex'] / r['n']:.3e}   {top}")
ops = collections.Counter()
for src, ex, wf, smp in ins:
    ops[src.split()[1].split(".")[0] if src.startswith("@") else src.split()[0].split(".")[0]] += ex
print("opcode mix:", ", ".join(f"{k}:{v / total:.1%}" for k, v in ops.most_common(24)))

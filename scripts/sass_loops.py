#!/usr/bin/env python
"""List the loops (backward branches) of one kernel's SASS with an opcode histogram per loop body.
usage: cuobjdump -sass -fun <mangled> file.o | python scripts/sass_loops.py [min_len]"""
import collections
import re
import sys

min_len = int(sys.argv[1]) if len(sys.argv) > 1 else 40
ins = []
for line in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
print("instructions:", len(ins))
for i, (a, text) in enumerate(ins):
    m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)\s*$", text)
    if not m:
        continue
    tgt = int(m.group(1), 16)
    if tgt <= a and tgt in addr_index:
        j = addr_index[tgt]
        body = ins[j:i + 1]
        if len(body) < min_len:
            continue
        hist = collections.Counter()
        for _, t in body:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            op = t.split()[0]
            hist[op.split(".")[0]] += 1
        print(f"loop {tgt:#x}..{a:#x}: {len(body)} instructions")
        print("   ", ", ".join(f"{k}:{v}" for k, v in hist.most_common()))

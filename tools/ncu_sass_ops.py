#!/usr/bin/env python
"""Executed-instruction histogram by SASS opcode from `ncu --page source --csv --print-source cuda,sass` output."""
import csv, sys, collections
rows = csv.reader(open(sys.argv[1]))
ops = collections.Counter(); smp = collections.Counter(); lsb = collections.Counter()
top = []
for r in rows:
    if len(r) < 40 or r[0] != "" or not r[2].startswith("0x"):
        continue
    toks = r[3].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0]
    try:
        n = int(r[7]); s = int(r[6])
    except ValueError:
        continue
    ops[op] += n; smp[op] += s
    if op in ("LDL", "STL"):
        top.append((n, r[3].strip()[:60]))
tot = sum(ops.values())
print("total", tot)
for op, n in ops.most_common(28):
    print(f"{op:10s} {n:12d} {100*n/tot:5.1f}%  samples {smp[op]}")
print("local-memory ops executed:")
for n, t in sorted(top, reverse=True)[:12]:
    print("  ", n, t)

#!/usr/bin/env python
"""Summarise an ncu --page source --csv dump: top SASS instructions by executed count / stall samples."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = rows[1]
iS, iE, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
data = [(r[iS].strip(), int(r[iE] or 0), int(r[iW] or 0), k) for k, r in enumerate(rows[2:]) if len(r) > iW]
totE = sum(d[1] for d in data); totW = sum(d[2] for d in data)
print(f"total executed {totE}  samples {totW}")
ops = {}
for s, e, w, k in data:
    op = s.split()[0] if not s.startswith("@") else s.split()[1]
    op = op.split(".")[0]
    o = ops.setdefault(op, [0, 0]); o[0] += e; o[1] += w
print("-- by opcode")
for op, (e, w) in sorted(ops.items(), key=lambda x: -x[1][0])[:25]:
    print(f"{op:12s} exec {100*e/totE:5.1f}%  stall {100*w/max(totW,1):5.1f}%")
print("-- top by stall samples")
for s, e, w, k in sorted(data, key=lambda d: -d[2])[:top]:
    print(f"{k:5d} stall {100*w/max(totW,1):5.1f}% exec {100*e/totE:5.2f}%  {s[:100]}")

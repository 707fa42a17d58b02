#!/usr/bin/env python
"""Per-CUDA-source-line summary of an `ncu --page source --csv --print-source cuda,sass` dump:
   share of executed warp instructions and of stall samples per line."""
import csv, sys
path = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
rows = list(csv.reader(open(path)))
cur_file = ""
lines = []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; iE = hdr.index("Instructions Executed"); iW = hdr.index("Warp Stall Sampling (All Samples)"); continue
    if hdr is None or len(r) <= iE: continue
    if r[0]:   # a CUDA line with aggregated metrics
        try: lines.append((cur_file, int(r[0]), r[1].strip(), int(r[iE] or 0), int(r[iW] or 0)))
        except ValueError: pass
totE = sum(l[3] for l in lines); totW = sum(l[4] for l in lines)
print(f"total executed {totE} samples {totW}")
for f, n, s, e, w in lines:
    if 100 * e / totE >= thr or 100 * w / max(totW, 1) >= thr:
        print(f"{f}:{n:<5d} exec {100*e/totE:5.1f}%  stall {100*w/max(totW,1):5.1f}%  {s[:100]}")

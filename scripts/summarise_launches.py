#!/usr/bin/env python
"""Filter an ncu launch list (--metrics gpu__time_duration.sum --csv) down to this library's kernels
(k_*, plus memsets are not kernels) and print each kernel's share of a hot-path step.

    python scripts/summarise_launches.py gpurun_out/launches_X.csv profiles/X_launches.csv > profiles/X_launches_summary.txt
"""
import collections
import csv
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
hdr = rows[0]
iK, iV, iG, iB, iM = (hdr.index(x) for x in ("Kernel Name", "Metric Value", "Grid Size", "Block Size", "Metric Name"))
mine = [r for r in rows[1:] if (r[iK].startswith("k_") or " k_" in r[iK].split("(")[0])]
with open(dst, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["launch", "kernel", "grid", "block", "metric", "value"])
    for r in mine:
        w.writerow([r[0], r[iK].split("(")[0].replace("void ", ""), r[iG], r[iB], r[iM], r[iV]])
names = sorted({r[iM] for r in mine})
for metric in names:
    agg = collections.OrderedDict()
    for r in mine:
        if r[iM] != metric:
            continue
        a = agg.setdefault(r[iK].split("(")[0].replace("void ", ""), [0, 0.0])
        a[0] += 1
        a[1] += float(r[iV].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"# {metric}: {len(mine)//max(1,len(names))} launches of this library's kernels "
          f"(the other launches in the capture are the torch kernels of the synthetic generator, outside the timed region)")
    print(f"{'kernel':40s} {'launches':>8s} {'per launch':>14s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:40s} {n:8d} {t / n:14.1f} {100 * t / tot:6.1f}%")

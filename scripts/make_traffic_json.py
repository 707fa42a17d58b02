#!/usr/bin/env python
"""profiles/traffic.json from the ncu CSVs of scripts/gpu_traffic.sh:
    python scripts/make_traffic_json.py r1h cfg5=gpurun_out/traffic_r1h_cfg5.csv cfg2=gpurun_out/traffic_r1h_cfg2.csv
Per workload and kernel group (bench.py's names): DRAM read + write bytes of ONE launch (last launch seen) and the
bytes per record at that workload's record count."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GROUP = {"k_coverage": "coverage", "k_coverage_tile": "coverage", "k_fine_accumulate_cluster": "accumulate", "k_split_bulk": "split", "k_accumulate": "accumulate", "k_ref_stats": "stats", "k_assign": "assign", "k_split": "split",
         "k_cutoffs_cluster": "cutoff", "k_cutoffs": "cutoff", "k_assign_reads": "assign", "k_fine_count": "accumulate",
         "k_fine_scan": "accumulate", "k_fine_split": "accumulate", "k_fine_accumulate": "accumulate"}
RECORDS = {"cfg2": 10_000_000, "cfg3": 100_000_000, "cfg4": 100_000_000, "cfg5": 1_000_000_000}
tag = sys.argv[1]
out = {"capture": tag, "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, bench.py at the full workload size"}
for spec in sys.argv[2:]:
    wl, path = spec.split("=")
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    iK, iM, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = {}
    for r in rows[1:]:
        k = r[iK].split("(")[0].replace("void ", "").split("<")[0]
        if k not in GROUP or not r[iM].startswith("dram__bytes"):
            continue
        d = per.setdefault(r[0], {"kernel": k, "bytes": 0.0})
        d["bytes"] += float(r[iV].replace(",", "")) * scale.get(r[iU], 1.0)
    # one hot-path step = the launches from one coverage kernel to the next; the last COMPLETE step counts, every launch of it
    # (k_fine_accumulate runs three times per step: packed, wide, cluster)
    order = [d for _, d in sorted(per.items(), key=lambda x: int(x[0]))]
    starts = [i for i, d in enumerate(order) if GROUP[d["kernel"]] == "coverage"]
    if len(starts) >= 2:
        order = order[starts[-2]:starts[-1]]
    last = {}
    for d in order:
        last[d["kernel"]] = last.get(d["kernel"], 0.0) + d["bytes"]
    res = {}
    for k, b in last.items():                                       # a group is the sum of its kernels (one launch each per step)
        g = res.setdefault(GROUP[k], {"kernel": [], "dram_bytes_per_launch": 0.0})
        g["kernel"].append(k); g["dram_bytes_per_launch"] += b
    for g in res.values():
        g["kernel"] = " + ".join(sorted(g["kernel"]))
        g["dram_bytes_per_record"] = g["dram_bytes_per_launch"] / RECORDS[wl]
    out[wl] = res
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))

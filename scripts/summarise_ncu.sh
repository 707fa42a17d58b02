#!/bin/bash
# usage: summarise_ncu.sh file.ncu-rep out_prefix   -> out_prefix_metrics.txt, _sass.txt, _lines.txt (what profiles/ keeps of a capture)
REP=$1; OUT=$2
bash scripts/ncu_metrics.sh "$REP" > ${OUT}_metrics.txt
ncu -i "$REP" --page source --csv > /tmp/_ncu_sass.csv 2>/dev/null && python scripts/ncu_sass.py /tmp/_ncu_sass.csv 40 > ${OUT}_sass.txt
ncu -i "$REP" --page source --csv --print-source cuda,sass > /tmp/_ncu_src.csv 2>/dev/null && python scripts/ncu_lines.py /tmp/_ncu_src.csv 1.0 > ${OUT}_lines.txt

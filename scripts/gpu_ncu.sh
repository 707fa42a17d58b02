#!/bin/bash
# full ncu captures of the kernels named in $KERNELS (regex each) on a 250 M-record cfg5 sample; tag $1; env passes through
# SKIP (default 1) launches of the kernel are skipped (warm-up steps), COUNT (default 1) are captured
mkdir -p gpurun_out
R=${1:-x}
for K in $KERNELS; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-1} -c ${COUNT:-1} -f -o gpurun_out/prof_${R}_$K \
      python bench.py --steps 1 --warmup 3 --records 250000000 --no-cpu-baseline --no-e2e > gpurun_out/prof_${R}_$K.log 2>&1
  echo "ncu $K rc=$?"
done

#!/bin/bash
# ncu launch list (per-launch durations) of the default bench; tag $1; env passes through
mkdir -p gpurun_out
R=${1:-x}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_${R}.log 2>&1
echo "launch list rc=$?"

#!/bin/bash
# ncu launch list of the default bench + full captures of the dominant kernels (1 GPU).
mkdir -p gpurun_out
R=${1:-r1}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_${R}.log 2>&1
echo "launch list rc=$?"
for K in ${KERNELS:-k_coverage k_assign k_accumulate}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_${R}_$K \
      python bench.py --steps 1 --warmup 3 --records 250000000 --no-cpu-baseline --no-e2e > gpurun_out/prof_${R}_$K.log 2>&1
  echo "ncu $K rc=$?"
done
ls -la gpurun_out

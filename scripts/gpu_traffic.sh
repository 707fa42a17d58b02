#!/bin/bash
# DRAM traffic per launch of every kernel of this library at the FULL workload sizes (one ncu pass set, 1 GPU):
# feeds profiles/traffic.json through scripts/make_traffic_json.py
mkdir -p gpurun_out
R=${1:-r1}
for WL in ${WORKLOADS:-cfg5 cfg2}; do
  timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:^k_ -c 60 --csv \
      --log-file gpurun_out/traffic_${R}_$WL.csv python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/traffic_${R}_$WL.log 2>&1
  echo "traffic $WL rc=$?"
done

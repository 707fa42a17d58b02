#!/bin/bash
# round-end evidence in one call (1 GPU): parity tests, smoke, bench lines (cfg5 default with e2e + CPU baseline + cli, cfg5 skip-bins,
# cfg2 / cfg3 / cfg4 on one GPU), launch list, DRAM traffic per kernel at full size, full ncu captures of the dominant kernels
mkdir -p gpurun_out
R=${1:-r2z}
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
fi
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${R}_cfg5.json 2> gpurun_out/bench_${R}_cfg5.err; echo "bench cfg5 rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --bins skip --no-cpu-baseline --no-e2e > gpurun_out/bench_${R}_cfg5_skipbins.json 2> gpurun_out/bench_${R}_cfg5_skipbins.err; echo "bench cfg5 skip rc=$?"
for W in cfg2 cfg3 cfg4; do
  timeout 600 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${R}_$W.json 2> gpurun_out/bench_${R}_$W.err; echo "bench $W rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_${R}.log 2>&1; echo "launch list rc=$?"
WORKLOADS="cfg5 cfg2" bash scripts/gpu_traffic.sh $R
# full captures (250 M-record cfg5 sample): the fourth k_fine_accumulate launch is the packed pass of the second step
# (the .ncu-rep files are summarised here and deleted: five of them exceed what gpurun copies back)
for K in k_coverage_tile k_assign_reads k_split k_fine_split k_fine_accumulate; do
  S=1; [ $K = k_fine_accumulate ] && S=3
  SKIP=$S KERNELS="$K" bash scripts/gpu_ncu.sh $R
  bash scripts/summarise_ncu.sh gpurun_out/prof_${R}_$K.ncu-rep gpurun_out/${R}_$K
  rm -f gpurun_out/prof_${R}_$K.ncu-rep
done

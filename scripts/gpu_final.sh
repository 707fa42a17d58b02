#!/bin/bash
# round-end evidence in one call (1 GPU): parity tests, smoke, bench lines (cfg5 default with e2e + CPU baseline, cfg5 skip-bins,
# cfg2), launch list, DRAM traffic per kernel at full size, full ncu captures of the three dominant kernels
mkdir -p gpurun_out
R=${1:-r1z}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${R}_cfg5.json 2> gpurun_out/bench_${R}_cfg5.err; echo "bench cfg5 rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --bins skip --no-cpu-baseline > gpurun_out/bench_${R}_cfg5_skipbins.json 2> gpurun_out/bench_${R}_cfg5_skipbins.err; echo "bench cfg5 skip rc=$?"
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 > gpurun_out/bench_${R}_cfg2.json 2> gpurun_out/bench_${R}_cfg2.err; echo "bench cfg2 rc=$?"
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_${R}_reference.json 2> gpurun_out/bench_${R}_reference.err; echo "bench reference rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_${R}.log 2>&1; echo "launch list rc=$?"
WORKLOADS="cfg5 cfg2" bash scripts/gpu_traffic.sh $R
KERNELS="k_coverage k_fine_accumulate k_assign_reads" bash scripts/gpu_ncu.sh $R

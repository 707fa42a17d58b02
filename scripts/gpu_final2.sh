#!/bin/bash
# refresh after the last kernel change of the round (the packed accumulate's fast path): tests, smoke, the cfg5 lines, launch list,
# traffic, one full capture of the packed k_fine_accumulate
mkdir -p gpurun_out
R=${1:-r2zz}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${R}_cfg5.json 2> gpurun_out/bench_${R}_cfg5.err; echo "bench cfg5 rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --bins skip --no-cpu-baseline --no-e2e > gpurun_out/bench_${R}_cfg5_skipbins.json 2> gpurun_out/bench_${R}_cfg5_skipbins.err; echo "bench cfg5 skip rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_${R}.log 2>&1; echo "launch list rc=$?"
WORKLOADS="cfg5" bash scripts/gpu_traffic.sh $R
SKIP=3 KERNELS="k_fine_accumulate" bash scripts/gpu_ncu.sh $R
bash scripts/summarise_ncu.sh gpurun_out/prof_${R}_k_fine_accumulate.ncu-rep gpurun_out/${R}_k_fine_accumulate
rm -f gpurun_out/prof_${R}_k_fine_accumulate.ncu-rep

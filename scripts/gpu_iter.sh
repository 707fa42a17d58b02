#!/bin/bash
# quick iteration: parity tests, default bench, optional ncu captures (KERNELS="k_a k_b", tag $1)
mkdir -p gpurun_out
R=${1:-it}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench cfg2 rc=$?"
for K in $KERNELS; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_${R}_$K \
      python bench.py --steps 1 --warmup 3 --records 250000000 --no-cpu-baseline --no-e2e > gpurun_out/prof_${R}_$K.log 2>&1
  echo "ncu $K rc=$?"
done

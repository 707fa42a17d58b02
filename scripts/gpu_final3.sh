#!/bin/bash
# last call of the round: the full GPU suite on the final tree + captures of two kernels the earlier calls did not cover
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
for K in k_fine_count k_cut_sort_cluster; do
  SKIP=1 KERNELS="$K" bash scripts/gpu_ncu.sh r2zz
  bash scripts/summarise_ncu.sh gpurun_out/prof_r2zz_$K.ncu-rep gpurun_out/r2zz_$K
  rm -f gpurun_out/prof_r2zz_$K.ncu-rep
done

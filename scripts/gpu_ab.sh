#!/bin/bash
# A/B of kernel variants selected by environment variables: parity tests (unless SKIP_TESTS=1), then the default bench per variant.
# usage: gpu_ab.sh "VAR=a VAR2=b" "VAR=c" ...   (each argument is one variant's environment; "" = defaults)
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
fi
i=0
for V in "$@"; do
  i=$((i+1))
  env $V timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e $BENCH_ARGS > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err
  echo "variant $i [$V] rc=$?"; tail -2 gpurun_out/ab_$i.err
  python -c "
import json
d=json.loads(open('gpurun_out/ab_$i.json').read().strip().splitlines()[-1]); print('  ', round(d['ms_per_step'],2),'ms', round(d['value']/1e9,2),'G rec/s', {k:round(v,2) for k,v in d['roofline']['pipeline']['kernel_ms'].items()}, d['result'])"
done

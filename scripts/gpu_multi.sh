#!/bin/bash
# needs gpurun --gpus N: multi-GPU parity test + the bench at every N in $1 (sharded_equals_single printed per run)
mkdir -p gpurun_out
GPUS=${1:-2}
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -5 gpurun_out/pytest_multi.log
fi
for N in $GPUS; do
  if [ $N -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline $BENCH_ARGS > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  else
    SLIMM_BENCH_PHASES=$PHASES timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 3 --warmup 3 $BENCH_ARGS > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  fi
  echo "bench N=$N rc=$?"; grep -E "phases|Error|error" gpurun_out/scale_$N.err | tail -3
  python -c "
import json,sys
d=json.loads(open('gpurun_out/scale_$N.json').read().strip().splitlines()[-1]); print('N=$N', round(d['ms_per_step'],2),'ms', round(d['value']/1e9,2),'G rec/s', 'e2e', d['e2e'] and round(d['e2e']['value']/1e9,2), {k:round(v,2) for k,v in d['roofline']['pipeline']['kernel_ms'].items()}, 'equal:', d.get('sharded_equals_single'), d.get('differing_fields'), d['result'])"
done

#!/bin/bash
# needs gpurun --gpus 8: cfg5 at N = 8, 4, 2, 1 (sharded_equals_single in every N > 1 line), then cfg3 and cfg4 at N = 8
mkdir -p gpurun_out
run() {  # workload N tag extra-args
  local W=$1 N=$2 T=$3; shift 3
  if [ $N -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --workload $W --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${T}.json 2> gpurun_out/${T}.err
  else
    SLIMM_BENCH_PHASES=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$((10+N)) bench.py --gpus $N --workload $W --steps 5 --warmup 3 "$@" > gpurun_out/${T}.json 2> gpurun_out/${T}.err
  fi
  echo "$T rc=$?"; grep -E "phases" gpurun_out/${T}.err | tail -1
  python -c "
import json
d=json.loads(open('gpurun_out/${T}.json').read().strip().splitlines()[-1]); print('$T', round(d['ms_per_step'],3),'ms', round(d['value']/1e9,2),'G rec/s', 'e2e', d['e2e'] and round(d['e2e']['value']/1e9,2), d['e2e'] and round(d['e2e'].get('h2d_probe_GBps_this_rank',0),1), {k:round(v,2) for k,v in d['roofline']['pipeline']['kernel_ms'].items()}, 'equal:', d.get('sharded_equals_single'), d.get('differing_fields'), d['result'])"
}
for N in ${NS:-8 4 2 1}; do run cfg5 $N scale_cfg5_$N; done
run cfg3 8 scale_cfg3_8
run cfg4 8 scale_cfg4_8
run cfg3 1 scale_cfg3_1
run cfg4 1 scale_cfg4_1
run cfg2 1 scale_cfg2_1

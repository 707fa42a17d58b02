#!/bin/bash
# needs gpurun --gpus 8: cfg5 at N = 8 and 4 (bench lines with per-rank kernel times and exchange phases)
mkdir -p gpurun_out
R=${1:-r2t}
for N in 8 4; do
  SLIMM_BENCH_PHASES=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_${R}_cfg5_$N.json 2> gpurun_out/scale_${R}_cfg5_$N.err
  echo "N=$N rc=$?"
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/scale_${R}_cfg5_$N.json') if l.startswith('{')][-1]); print('N=$N', round(d['ms_per_step'],3),'ms', round(d['value']/1e9,2),'G rec/s', 'e2e', d['e2e'] and round(d['e2e']['value']/1e9,2), d['e2e'] and d['e2e'].get('h2d_probe_GBps_per_rank'), {k:round(v,2) for k,v in d['roofline']['pipeline']['kernel_ms'].items()}, 'equal:', d.get('sharded_equals_single'), d.get('differing_fields')); print(d.get('exchange_phases_ms'))"
done

#!/bin/bash
# needs gpurun --gpus 8: cfg5, cfg3 and cfg4 at N = 8 only (bench lines with per-rank kernel times and exchange phases)
mkdir -p gpurun_out
R=${1:-r2}
for W in cfg5 cfg3 cfg4; do
  SLIMM_BENCH_PHASES=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --workload $W --steps 5 --warmup 3 > gpurun_out/scale_${R}_${W}_8.json 2> gpurun_out/scale_${R}_${W}_8.err
  echo "$W rc=$?"; grep -E "phases" gpurun_out/scale_${R}_${W}_8.err | tail -1
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/scale_${R}_${W}_8.json') if l.startswith('{')][-1]); print('$W', round(d['ms_per_step'],3),'ms', round(d['value']/1e9,2),'G rec/s', 'e2e', d['e2e'] and round(d['e2e']['value']/1e9,2), {k:round(v,2) for k,v in d['roofline']['pipeline']['kernel_ms'].items()}, 'equal:', d.get('sharded_equals_single'), d.get('differing_fields')); print(d.get('per_rank_kernel_ms')); print(d.get('exchange_phases_ms'))"
done

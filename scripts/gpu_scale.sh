#!/bin/bash
# needs gpurun --gpus 8: the bench at N = 8 and 4 (peer-to-peer split), phases on stderr
mkdir -p gpurun_out
for N in ${NS:-8 4}; do
  SLIMM_BENCH_PHASES=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 3 --warmup 3 --no-e2e > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  echo "bench N=$N rc=$?"; grep phases gpurun_out/scale_$N.err
  python -c "
import json
d=json.loads(open('gpurun_out/scale_$N.json').read().strip().splitlines()[-1]); print('N=$N', round(d['ms_per_step'],2),'ms', round(d['value']/1e9,2),'G rec/s', d['config'].get('exchange'), {k:round(v,2) for k,v in d['roofline']['pipeline']['kernel_ms'].items()})"
done

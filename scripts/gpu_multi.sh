#!/bin/bash
# needs gpurun --gpus N: multi-GPU parity test + scaling of the bench (N=1..GPUS)
mkdir -p gpurun_out
GPUS=${1:-2}
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -15 gpurun_out/pytest_multi.log
for N in $GPUS; do
  if [ $N -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  fi
  echo "bench N=$N rc=$?"; tail -3 gpurun_out/scale_$N.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/scale_$N.json').read().strip().splitlines()[-1]); print('N=$N', round(d['ms_per_step'],2),'ms', round(d['value']/1e9,2),'G rec/s', {k:round(v,2) for k,v in d['roofline']['pipeline']['kernel_ms'].items()})"
done
SLIMM_BENCH_P2P=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $GPUS --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $GPUS --steps 3 --warmup 3 --no-e2e > gpurun_out/scale_${GPUS}_nccl.json 2> gpurun_out/scale_${GPUS}_nccl.err
python -c "
import json
d=json.loads(open('gpurun_out/scale_${GPUS}_nccl.json').read().strip().splitlines()[-1]); print('nccl a2a N=$GPUS', round(d['ms_per_step'],2),'ms', round(d['value']/1e9,2),'G rec/s', {k:round(v,2) for k,v in d['roofline']['pipeline']['kernel_ms'].items()})"

#!/bin/bash
# decoder A/B on the GPU box's host cores: a 10 M-record cfg5-shaped SAM, old vs new front end, 8 and 16 threads
mkdir -p gpurun_out /tmp/dec
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from slimm_b200 import synth
rng = np.random.default_rng(1)
tax, accs = synth.make_taxonomy(50000)
contigs = synth.make_contigs(50000, rng, accs)
rec = synth.make_records(contigs, 10_000_000, np.random.default_rng(2), multi_frac=0.2)
synth.write_sam_for_records('/tmp/dec/in.sam', contigs, rec)
PY
g++ -O2 -std=c++17 -pthread -I slimm_b200/csrc/frontend scripts/micro/decode_bench.cpp -lz -o /tmp/dec/new
g++ -O2 -std=c++17 -pthread -I scripts/micro/_old_fe scripts/micro/decode_bench.cpp -lz -o /tmp/dec/old
nproc
for t in 8 16; do for v in old new; do echo "$v $t threads: $(/tmp/dec/$v /tmp/dec/in.sam $t | tail -1)"; done; done | tee gpurun_out/decode_ab.txt

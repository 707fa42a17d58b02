#!/bin/bash
# the command line's phase breakdown (-v) on a cfg5-shaped 3.4 M-record SAM: what the fixed cost of a run is made of
mkdir -p gpurun_out /tmp/clip/out
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from slimm_b200 import synth, sldb
rng = np.random.default_rng(20260101)
tax, accs = synth.make_taxonomy(50000)
contigs = synth.make_contigs(50000, rng, accs)
rec = synth.make_records(contigs, 3_375_000, np.random.default_rng(20260102), multi_frac=0.2)
synth.write_sam_for_records('/tmp/clip/in.sam', contigs, rec)
sldb.write_sldb(synth.database_for(tax), '/tmp/clip/db.sldb')
PY
for i in 1 2 3; do
  echo "run $i"; ( time slimm_b200/bin/slimm -v -w 100 -cc 0.95 -o /tmp/clip/out/ /tmp/clip/db.sldb /tmp/clip/in.sam ) 2>&1 | grep -E "phases|decode:|real|GPU context|database|rror"
done | tee gpurun_out/cli_phases.txt

#!/bin/bash
# per-kernel A/B: ncu durations of the kernels matching $1 (regex) for every environment variant that follows
# usage: gpu_kab.sh 'k_fine_split|k_split' "VAR=a" "VAR=b" ...     (RECORDS: sample size, default 500 M records of cfg5)
mkdir -p gpurun_out
K=$1; shift
i=0
for V in "$@"; do
  i=$((i+1))
  env $V timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:$K" -c ${COUNT:-24} --csv --log-file gpurun_out/kab_$i.csv \
      python bench.py --steps 2 --warmup 2 --records ${RECORDS:-500000000} --no-cpu-baseline --no-e2e $BENCH_ARGS > gpurun_out/kab_$i.log 2>&1
  echo "variant $i [$V] rc=$?"
  python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/kab_$i.csv') if l.startswith('"')))
h = rows[0]; iK = h.index('Kernel Name'); iV = h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[iK].split('(')[0].replace('void ', ''), []).append(float(r[iV].replace(',', '')) / 1e3)
for k, v in agg.items():
    v = v[len(v) // 2:]          # second half: past the warm-up
    print(f'   {k:45s} {len(v):3d} launches  {sum(v) / len(v):9.1f} us  (min {min(v):.1f})')
PY
done

for S in 256 512 1024 1536 2048 3072; do
  SLIMM_ACC_SHAPE=$S timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['pipeline']['kernel_ms']; print('$S', round(d['ms_per_step'],2), 'accumulate', round(k['accumulate'],3), 'memset', round(k['sort'],3))"
done

timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for S in 6 8; do
  SLIMM_COV_CTAS=$S timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['pipeline']['kernel_ms']; print('cov_ctas $S', round(d['ms_per_step'],2), {a:round(b,2) for a,b in k.items()})"
done

timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for v in 1 0; do
 for mb in 3 4; do
  echo "LEAN=$v BORDER_MB=$mb"
  GCM_B200_LEAN_BORDER=$v GCM_B200_BORDER_MB=$mb GCM_B200_CONCURRENT=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['roofline'].get('kernel_ms_share'))
"
 done
done
echo CONCURRENT
GCM_B200_LEAN_BORDER=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | cut -c1-600

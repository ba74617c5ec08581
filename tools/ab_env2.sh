# like ab_env.sh but concurrent streams on (product default) and 10 steps
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
for envs in "$@"; do
  echo "== $envs"
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_step'],3), d['gpu_launches'], {k: round(v,2) for k,v in d['roofline'].get('kernel_ms_share').items()})
"
done

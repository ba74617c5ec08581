for envs in "$@"; do
  echo "== $envs"
  env $envs GCM_B200_CONCURRENT=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_step'],3), {k: round(v,2) for k,v in d['roofline'].get('kernel_ms_share').items()})
"
done

for envs in "$@"; do
  echo "== $envs"
  env $envs timeout 300 python bench.py --size 108 --zones 1 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_step'],4), {k: round(v,2) for k,v in d['roofline'].get('kernel_ms_share').items()})
"
done

# usage: r2_ab.sh "<ENV=...>" ... -- one short bench line per environment setting
for e in "$@"; do
  env $e timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$e', 'ms/step', round(d['ms_per_step'],4), 'share', {k: round(v,3) for k,v in d['roofline']['kernel_ms_share'].items()})"
done

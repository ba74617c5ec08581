# usage: r2_round.sh <tag> [bench args...]  -- parity tests, then serial + concurrent bench lines
TAG=$1; shift
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_$TAG.log 2>&1; tail -3 gpurun_out/pytest_$TAG.log
GCM_B200_CONCURRENT=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>gpurun_out/bench_serial_$TAG.err | tail -1 > gpurun_out/bench_serial_$TAG.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_serial_$TAG.json')); print('serial', d['ms_per_step'], d['roofline'].get('kernel_ms_share'), d['roofline']['frac'])
except Exception as e: print('serial failed', e); print(open('gpurun_out/bench_serial_$TAG.err').read()[-2000:])
PY
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" 2>gpurun_out/bench_$TAG.err | tail -1 > gpurun_out/bench_$TAG.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$TAG.json')); print('concurrent', d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['step_frac'], d['e2e'])
except Exception as e: print('bench failed', e); print(open('gpurun_out/bench_$TAG.err').read()[-2000:])
PY

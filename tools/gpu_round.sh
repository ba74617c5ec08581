# usage: gpu_round.sh <tag> <kernel-regex> [bench args...]  -- parity tests, bench line, one ncu capture
TAG=$1; KRE=$2; shift 2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
GCM_B200_CONCURRENT=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('serial', d['ms_per_step'], d['roofline'].get('kernel_ms_share'))
"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/bench_$TAG.json; python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print('concurrent', d['ms_per_step'], d['value'], d['roofline']['frac'])"
python bench.py --size 100 --zones 1 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KRE --launch-skip 3 -c 1 -o gpurun_out/prof_$TAG -f python bench.py --size 100 --zones 1 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep

# bench lines (serial + concurrent) then ncu launch list
TAG=$1; shift
GCM_B200_CONCURRENT=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>gpurun_out/bench_serial_$TAG.err | tail -1 > gpurun_out/bench_serial_$TAG.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_serial_$TAG.json')); print('serial', d['ms_per_step'], d['roofline'].get('kernel_ms_share'), d['roofline']['frac'], d['roofline']['coverage'])
except Exception as e: print('serial failed', e); print(open('gpurun_out/bench_serial_$TAG.err').read()[-2000:])
PY
timeout 600 python bench.py --steps 20 --warmup 3 "$@" 2>gpurun_out/bench_$TAG.err | tail -1 > gpurun_out/bench_$TAG.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$TAG.json')); print('concurrent', d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['step_frac'], d['e2e'], d['cpu_baseline'], d['state_digest'], d['config']['setup_s'])
except Exception as e: print('bench failed', e); print(open('gpurun_out/bench_$TAG.err').read()[-2000:])
PY
bash tools/r2_ncu_list.sh $TAG "$@"

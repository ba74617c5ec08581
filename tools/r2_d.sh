# 1 GPU: whole parity suite, host-detector A/B on the contact cases, full bench line with the extra configs
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_d.log 2>&1; tail -12 gpurun_out/pytest_d.log
GCM_B200_DEVICE_COLLISIONS=0 timeout 600 python -m pytest tests -m gpu -q -k "contact or friction or virtual" > gpurun_out/pytest_d_hostcoll.log 2>&1; tail -2 gpurun_out/pytest_d_hostcoll.log
timeout 900 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_d.err | tail -1 > gpurun_out/bench_d.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_d.json')); print('bench', d['ms_per_step'], 'host issue', d['host_issue_ms_per_step'], d['value'], d['roofline']['step_frac'], d['e2e'], d['cpu_baseline'], d['state_digest'])
    for k,v in d['other_configs'].items(): print(k, {a: v[a] for a in ('ms_per_step','value','step_frac')}, v.get('contact_nodes'), v.get('kernel_ms_per_5_steps_serialised'))
except Exception as e: print('bench failed', e); print(open('gpurun_out/bench_d.err').read()[-3000:])
PY

# 1 GPU: whole parity suite, register-budget A/B of the pattern kernel, full bench line with the extra configs
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_c.log 2>&1; tail -15 gpurun_out/pytest_c.log
bash tools/r2_ab.sh GCM_B200_INNER_MB=1 GCM_B200_INNER_MB=4 GCM_B200_INNER_MB=5 GCM_B200_INNER_MB=6
timeout 900 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_c.err | tail -1 > gpurun_out/bench_c.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_c.json')); print('bench', d['ms_per_step'], 'host issue', d['host_issue_ms_per_step'], d['value'], d['roofline']['step_frac'], d['e2e'], d['cpu_baseline'], d['state_digest'])
    for k,v in d['other_configs'].items(): print(k, {a: v[a] for a in ('ms_per_step','value','step_frac')}, v.get('contact_nodes'), v.get('kernel_ms_per_5_steps_serialised'))
except Exception as e: print('bench failed', e); print(open('gpurun_out/bench_c.err').read()[-3000:])
PY

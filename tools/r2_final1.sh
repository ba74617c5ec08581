# 1 GPU, final state: parity suite, bench line (with extra configs), launch list, ncu --set full of one launch of each hot kernel
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_final.log 2>&1; tail -4 gpurun_out/pytest_final.log
timeout 900 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_final_n1.err | tail -1 > gpurun_out/bench_final_n1.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_final_n1.json')); print('bench', d['ms_per_step'], d['value'], d['roofline']['step_frac'], d['e2e']['value'], d['state_digest'])
    for k,v in d['other_configs'].items(): print(k, {a: v[a] for a in ('ms_per_step','value','step_frac')}, v.get('contact_nodes'), v.get('kernel_ms_per_5_steps_serialised'))
except Exception as e: print('bench failed', e); print(open('gpurun_out/bench_final_n1.err').read()[-3000:])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_final_reference.json; cut -c1-300 gpurun_out/bench_final_reference.json
bash tools/r2_ncu_list.sh final --no-extra-configs | tail -25
GCM_B200_CONCURRENT=0 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"stage_border_fast_kernel|stage_inner_pattern_kernel" --launch-skip 6 -c 6 -o gpurun_out/prof_final -f \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra-configs > gpurun_out/ncu_final.log 2>&1
ls -la gpurun_out/prof_final.ncu-rep

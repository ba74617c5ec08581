# 1 GPU, final state (short form): bench line with extra configs, ncu --set full of one launch of each hot kernel
timeout 900 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_final_n1.err | tail -1 > gpurun_out/bench_final_n1.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_final_n1.json')); print('bench', d['ms_per_step'], d['value'], d['roofline']['step_frac'], d['e2e']['value'], d['state_digest'], d['roofline']['kernel_sha'])
    for k,v in d['other_configs'].items(): print(k, {a: v[a] for a in ('ms_per_step','value','step_frac')}, v.get('contact_nodes'))
except Exception as e: print('bench failed', e); print(open('gpurun_out/bench_final_n1.err').read()[-3000:])
PY
bash tools/r2_ncu_list.sh final --no-extra-configs | tail -8
GCM_B200_CONCURRENT=0 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"stage_border_fast_kernel|stage_inner_pattern_kernel" --launch-skip 6 -c 6 -o gpurun_out/prof_final -f \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra-configs > gpurun_out/ncu_final.log 2>&1
ls -la gpurun_out/prof_final.ncu-rep

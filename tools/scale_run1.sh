timeout 600 python bench.py --gpus 1 --steps 40 --warmup 5 --no-extra-configs > gpurun_out/bench_final_n1_s40.log 2>&1
tail -1 gpurun_out/bench_final_n1_s40.log > gpurun_out/bench_final_n1_s40.json
python -c "
import json; d=json.load(open('gpurun_out/bench_final_n1_s40.json')); print(1, d['ms_per_step'], d['value'], d['e2e']['value'], d['state_digest'], d['roofline']['traffic'], d['roofline']['traffic_frac'], d['roofline']['frac_kernel_bytes'])" || tail -20 gpurun_out/bench_final_n1_s40.log

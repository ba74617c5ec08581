# 8-GPU box: bench line at N=8 (peer stores)
N=8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline ${EXTRA} > gpurun_out/bench_n${N}.log 2>&1
tail -1 gpurun_out/bench_n${N}.log > gpurun_out/bench_n${N}.json
python -c "
import json; d=json.load(open('gpurun_out/bench_n${N}.json')); print('N=$N', d['ms_per_step'], 'host issue', d['host_issue_ms_per_step'], d['value'], d['roofline']['kernel_ms_share'], d['e2e']['value'], d['state_digest'], d['gpu_launches'])" || tail -20 gpurun_out/bench_n${N}.log

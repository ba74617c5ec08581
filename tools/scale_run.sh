# usage: scale_run.sh N -- the bench line at N GPUs as the driver launches it (defaults but for the step count)
N=$1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/bench_final_n$N.log 2>&1
tail -1 gpurun_out/bench_final_n$N.log > gpurun_out/bench_final_n$N.json
python -c "
import json; d=json.load(open('gpurun_out/bench_final_n$N.json')); print($N, d['ms_per_step'], d['value'], d['roofline']['kernel_ms_share'], d['e2e']['value'], d['state_digest'], d['host_issue_ms_per_step'], d['roofline']['coverage'])" || tail -20 gpurun_out/bench_final_n$N.log

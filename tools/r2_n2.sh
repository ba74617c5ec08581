# 2 GPUs: cross-rank parity tests, then bench lines: overlapped exchange on / off
timeout 900 python -m pytest tests/test_multiprocess.py -m gpu -x -q 2>&1 | tail -3
for D in 1 0; do
GCM_B200_OVERLAP=$D timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$D bench.py --gpus 2 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2_ov$D.log 2>&1
tail -1 gpurun_out/bench_n2_ov$D.log > gpurun_out/bench_n2_ov$D.json
python -c "
import json; d=json.load(open('gpurun_out/bench_n2_ov$D.json')); print('overlap=$D', d['ms_per_step'], d['value'], d['roofline']['kernel_ms_share'], d['e2e']['value'], d['state_digest'], d['roofline']['coverage'])" || tail -20 gpurun_out/bench_n2_ov$D.log
done

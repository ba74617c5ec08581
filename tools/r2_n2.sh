# 2 GPUs: cross-rank parity test, then bench lines: copy-engine exchange, push-kernel exchange
timeout 600 python -m pytest tests/test_multiprocess.py -m gpu -x -q 2>&1 | tail -3
for D in 1 0; do
GCM_B200_P2P_DMA=$D timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$D bench.py --gpus 2 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2_dma$D.log 2>&1
tail -1 gpurun_out/bench_n2_dma$D.log > gpurun_out/bench_n2_dma$D.json
python -c "
import json; d=json.load(open('gpurun_out/bench_n2_dma$D.json')); print('dma=$D', d['ms_per_step'], d['value'], d['roofline']['kernel_ms_share'], d['e2e']['value'], d['state_digest'], d['config']['layers'])" || tail -20 gpurun_out/bench_n2_dma$D.log
done

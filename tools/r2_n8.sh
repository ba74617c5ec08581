# 8-GPU box: bench line at N=8 (peer stores)
for cfg in "8 1"; do set -- $cfg; N=$1; P=$2
GCM_B200_P2P=$P timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline ${EXTRA} > gpurun_out/bench_n${N}_p2p$P.log 2>&1
tail -1 gpurun_out/bench_n${N}_p2p$P.log > gpurun_out/bench_n${N}_p2p$P.json
python -c "
import json; d=json.load(open('gpurun_out/bench_n${N}_p2p$P.json')); print('N=$N p2p=$P', d['ms_per_step'], d['value'], d['roofline']['kernel_ms_share'], d['e2e']['value'], d['state_digest'], d['gpu_launches'], d['roofline']['coverage'])" || tail -20 gpurun_out/bench_n${N}_p2p$P.log
done

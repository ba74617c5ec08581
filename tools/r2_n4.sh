# 4 GPUs: small-mesh digest check (peer stores vs 1 GPU), the cross-rank parity test, then the 216^3 bench line
timeout 200 python bench.py --size 72 --zones 8 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1 small', d['ms_per_step'], d['state_digest'])"
GCM_B200_P2P_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --size 72 --zones 8 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/dbg_n4.log 2>&1
grep "mapping check\|out of range" gpurun_out/dbg_n4.log | head -4
tail -1 gpurun_out/dbg_n4.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=4 small', d['ms_per_step'], d['state_digest'], d['roofline']['coverage'])" || grep -n "GCMExc" gpurun_out/dbg_n4.log | head
timeout 600 python -m pytest tests/test_multiprocess.py -m gpu -x -q 2>&1 | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n4.log 2>&1
tail -1 gpurun_out/bench_n4.log > gpurun_out/bench_n4.json
python -c "
import json; d=json.load(open('gpurun_out/bench_n4.json')); print('N=4', d['ms_per_step'], d['value'], d['roofline']['kernel_ms_share'], d['e2e']['value'], d['state_digest'], d['gpu_launches'], d['roofline']['coverage'])" || tail -20 gpurun_out/bench_n4.log

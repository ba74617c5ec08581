# 4 GPUs, small mesh: peer-store exchange against the single-GPU digest
timeout 200 python bench.py --size 72 --zones 8 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1', d['ms_per_step'], d['state_digest'])"
for P in 1 0; do
GCM_B200_P2P_DEBUG=1 GCM_B200_P2P=$P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2954$P bench.py --gpus 4 --size 72 --zones 8 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/dbg_n4_p2p$P.log 2>&1
grep "gcm_b200 p2p" gpurun_out/dbg_n4_p2p$P.log | head -20
tail -1 gpurun_out/dbg_n4_p2p$P.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=4 p2p=$P', d['ms_per_step'], d['state_digest'], d['roofline']['coverage'])" || grep -n "GCMExc" gpurun_out/dbg_n4_p2p$P.log | head
done

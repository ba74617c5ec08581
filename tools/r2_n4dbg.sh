for ONLY in 1 2 3; do
GCM_B200_P2P_ONLY=$ONLY GCM_B200_P2P_DEBUG=1 GCM_B200_P2P=1 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2954$ONLY bench.py --gpus 4 --size 72 --zones 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/dbg_n4_only$ONLY.log 2>&1
echo "== only rank $ONLY pushes"; grep "seq 1 \|seq 2 push" gpurun_out/dbg_n4_only$ONLY.log | head -12
done

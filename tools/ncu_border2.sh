python bench.py --size 100 --zones 1 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain_b2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stage_border_fused_kernel --launch-skip 3 -c 2 -o gpurun_out/prof_r1s_border_fused -f python bench.py --size 100 --zones 1 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_b2.log 2>&1
ls -la gpurun_out/prof_r1s_border_fused.ncu-rep

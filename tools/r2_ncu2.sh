# ncu --set full of the first launches of the border and inner kernels (after a plain run exits 0)
GCM_B200_CONCURRENT=0 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra-configs > gpurun_out/plain_r2b.log 2>&1 && \
GCM_B200_CONCURRENT=0 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"stage_border_fast_kernel|stage_inner_pattern_kernel" --launch-skip 6 -c 6 -o gpurun_out/prof_r2b -f \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra-configs > gpurun_out/ncu_r2b.log 2>&1
ls -la gpurun_out/prof_r2b.ncu-rep; tail -3 gpurun_out/ncu_r2b.log

python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain_list.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1n.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1
wc -l gpurun_out/launches_r1n.csv

# usage: ncu_full.sh <tag> <kernel regex> <launch-skip>
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain_full.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 --launch-skip $3 -c 1 -o gpurun_out/prof_$1 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/prof_$1.ncu-rep

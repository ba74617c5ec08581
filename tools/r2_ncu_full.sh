# usage: r2_ncu_full.sh <tag> <kernel-regex> <skip> <count> [bench args] -- ncu --set full of selected launches (after a plain run exits 0)
TAG=$1; KRE=$2; SKIP=$3; CNT=$4; shift 4
GCM_B200_CONCURRENT=0 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e "$@" > gpurun_out/plain_$TAG.log 2>&1 && \
GCM_B200_CONCURRENT=0 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$KRE" --launch-skip $SKIP -c $CNT -o gpurun_out/prof_$TAG -f \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ncu_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep; tail -3 gpurun_out/ncu_$TAG.log

#!/usr/bin/env bash
# usage: gpurun_retry.sh <logfile> <timeout> <command...>  -- retries while the pod answers "busy" (nothing is charged for those)
LOG=$1; TO=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > "$LOG" 2>&1
  if grep -q "status=transient" "$LOG"; then sleep 120; continue; fi
  break
done
tail -30 "$LOG"

#!/usr/bin/env bash
# usage: [GPUS=N] gpurun_retry.sh <logfile> <timeout> <command...>  -- retries while the pod answers "busy" (nothing is charged for those)
LOG=$1; TO=$2; shift 2
G=""; [ -n "$GPUS" ] && G="--gpus $GPUS"
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun $G --timeout $TO -- "$@" > "$LOG" 2>&1
  if grep -q "status=transient" "$LOG"; then sleep 120; continue; fi
  break
done
tail -40 "$LOG"

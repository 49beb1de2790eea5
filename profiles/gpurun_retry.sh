#!/bin/bash
# gpurun with retries on "no slot right now" (rc 3 / transient):  bash profiles/gpurun_retry.sh <timeout> '<command>' <logfile>
T=$1; CMD=$2; LOG=$3
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun ${GPURUN_ARGS:-} --timeout $T -- "$CMD" > $LOG 2>&1
  if grep -q "status=transient\|rc=3\|answers busy\|retry in a few minutes" $LOG && ! grep -q "status=ok" $LOG; then sleep 90; continue; fi
  break
done

#!/bin/bash
# usage: [GPUS=N] scripts/gpurun_retry.sh <logfile> <timeout_s> <command...>   (retries while the pod answers busy)
LOG=$1; shift; TMO=$1; shift
G=""; [ -n "$GPUS" ] && G="--gpus $GPUS"
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $G --timeout $TMO -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient" $LOG || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
echo "gpurun_retry done rc=$rc" >> $LOG

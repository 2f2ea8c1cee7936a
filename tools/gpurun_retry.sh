#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout> <command...>   — retries while the pod answers busy (exit code 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient" $log || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
echo "done rc=$rc" >> $log

#!/bin/bash
# usage: gpurun_retry.sh <timeout> <logfile> <command...>   -- retries while the pod answers "transient"
TO=$1; LOG=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  if ! grep -q "status=transient" $LOG; then exit 0; fi
  sleep 150
done
exit 3

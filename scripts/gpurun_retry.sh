#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <timeout> <command...> : retries while the pod answers busy/transient (nothing is charged then)
log=$1; shift; to=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|status=busy\|rc=None" $log && ! grep -q "status=ok" $log; then sleep 90; else break; fi
done

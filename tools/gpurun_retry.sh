#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE [gpurun args...] -- 'command'   : retries while the pod answers "transient" (nothing charged)
LOG=$1; shift
for attempt in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if grep -q "status=transient" "$LOG"; then sleep 200; continue; fi
  break
done
tail -3 "$LOG"

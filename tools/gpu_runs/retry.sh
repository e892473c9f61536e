#!/bin/bash
# usage: retry.sh [gpurun options] -- command : retries while the pod answers "transient / busy" (nothing charged)
for attempt in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | grep -v "^\[gpurun\] sending"
  if echo "$out" | grep -q "status=transient\|rc=3\|busy"; then sleep 150; continue; fi
  break
done

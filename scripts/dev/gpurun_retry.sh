#!/bin/bash
# scripts/dev/gpurun_retry.sh <timeout_s> <command...>: retries while the pod answers "busy" (exit 3)
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@"; rc=$?
  if [ $rc -ne 3 ] && ! grep -q '"status": "transient"' gpurun_out/.last_call.json 2>/dev/null; then exit $rc; fi
  sleep 90
done

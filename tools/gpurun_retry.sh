#!/bin/bash
# dev helper: retry a gpurun call while the pod answers "busy" (transient, nothing charged)
# usage: tools/gpurun_retry.sh <timeout_s> [--gpus N] -- '<command>'
T=$1; shift
for i in $(seq 1 24); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@"
  rc=$?
  if ! grep -q '"status": "transient"' gpurun_out/.last_call.json 2>/dev/null; then exit $rc; fi
  echo "[retry] busy (rc=$rc), attempt $i; sleeping 120 s" >&2
  sleep 120
done
exit 3

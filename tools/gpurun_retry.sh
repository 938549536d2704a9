#!/bin/bash
# Developer helper: retry a gpurun call while the pod answers "transient" (no slot free; nothing is charged).
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>' [extra gpurun flags...]
to=$1; cmd=$2; shift 2
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" --timeout "$to" -- "$cmd" 2>&1)
  echo "$out" | tail -4
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  break
done

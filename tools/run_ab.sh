#!/bin/bash
# Developer helper (GPU box): A/B of harness builds in product mode (mode 0) + one timeline. usage: tools/run_ab.sh <log> <binaries...>
L=$1; shift
: > $L
for b in "$@"; do
  echo "=== $b" >> $L
  timeout 100 $b time 4096 10000001 20 0 -1 1 1 0 2 2>&1 | grep -E "time\]|error|CUDA" >> $L
  timeout 100 $b time 32768 1250001 20 0 -1 1 1 0 2 2>&1 | grep -E "time\]|error|CUDA" >> $L
done
for b in "$@"; do
  echo "=== timeline $b" >> $L
  LRB_TIMELINE=1 timeout 100 $b time 4096 10000001 20 6 -1 1 1 0 2 2>&1 | grep -E "time\]|per CTA|error|CUDA|timeline|t\+" >> $L
done

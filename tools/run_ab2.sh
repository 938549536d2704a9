#!/bin/bash
# Developer helper (GPU box): product-mode A/B of harness builds, two rounds. usage: tools/run_ab2.sh <log> <binaries...>
L=$1; shift
: > $L
for round in 1 2; do
for b in "$@"; do
  echo "=== $b" >> $L
  timeout 100 $b time 4096 10000001 20 0 -1 1 1 0 3 2>&1 | grep -E "time\]|error|CUDA" >> $L
  timeout 100 $b time 32768 1250001 20 0 -1 1 1 0 2 2>&1 | grep -E "time\]|error|CUDA" >> $L
done
done

#!/bin/bash
# Developer helper (GPU box): one rank's scoring call of the N-GPU weak-scaling step (4096*N gathered users x 10M/N local
# rows) with different users-per-launch caps (tools/tc_check arg 9: pair tiles per launch = 74 / capdiv)
for cfg in "8192 5000001" "16384 2500001" "32768 1250001"; do
  for capdiv in 2 1; do
    echo "=== users/rows $cfg capdiv $capdiv"
    timeout 300 tools/tc_check time $cfg 20 0 -1 1 1 $capdiv 2 2>&1 | grep -E "time\]|error|CUDA"
  done
done

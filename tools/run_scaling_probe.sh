#!/bin/bash
# Developer helper (GPU box): one rank's scoring call of the N-GPU weak-scaling step (4096*N gathered users x 10M/N local
# rows) for different restart-cost weights of the work decomposition (tools/tc_check arg 11), A/B on one box
for cfg in "4096 10000001" "16384 2500001" "32768 1250001"; do
  for r0 in 0 150 300 600; do
    echo "=== users/rows $cfg restart-tiles $r0"
    timeout 300 tools/tc_check time $cfg 20 0 -1 1 1 0 2 $r0 2>&1 | grep -E "time\]|error|CUDA"
  done
done

#!/bin/bash
# Developer helper (GPU box): pipeline probes of the scoring kernel via tools/tc_check (see PROBE in score_topk_tc.cuh)
# usage: tools/run_probes.sh <rows> <which: 1 = bias+exclusion, 2 = no bias, 3 = both> [modes...]
rows=${1:-10000001}; which=${2:-2}; shift 2
modes=${@:-6 5 1 2 4}
for m in $modes; do
  echo "=== mode $m rows $rows ${LRB_SCORE_CTA_GROUP:+cta_group $LRB_SCORE_CTA_GROUP}"
  timeout 300 ${TC:-tools/tc_check} time 4096 $rows 20 $m -1 1 1 0 $which 2>&1 | grep -E "time\]|per CTA|error|CUDA"
done

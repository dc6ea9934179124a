#!/bin/bash
# env sweep of the cluster chunk / warps of k_rows_cl (2 CTAs per SM hypothesis)
for cw in "512 12" "512 8" "384 6" "320 5" "256 4" "256 5" "256 6" "192 4" "128 3" "128 4"; do
  set -- $cw
  r=$(AFB_ROWS_CHUNK=$1 AFB_ROWS_WARPS=$2 python bench.py --steps 5 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['step']['gather_ms'], d['roofline']['step']['element_ms'])")
  echo "chunk=$1 warps=$2 -> $r"
done

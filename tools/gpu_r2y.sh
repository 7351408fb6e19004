#!/bin/bash
# k_wide_tile hand-over tuning on the English-like dictionary
mkdir -p gpurun_out
TAG=${1:-r2y}
for SPEC in ${SWEEP:-3:768 2:1024 2:1536 2:2048 1:2048 1:4096 3:1536 4:512}; do
  IFS=: read MR HO <<< "$SPEC"
  ACGPU_WT_MIN_ROUNDS=$MR ACGPU_WT_HANDOVER=$HO timeout 300 python tools/bench_configs.py --configs 5 --scale 0.25 --steps 3 --warmup 2 --e2e-chars 1000000 > gpurun_out/${TAG}_$MR_$HO.jsonl 2>/dev/null
  python - <<PY
import json
for ln in open("gpurun_out/${TAG}_$MR_$HO.jsonl"):
    d = json.loads(ln)
    print("min_rounds $MR hand_over $HO  %-30s %8.2f ms %7.1f GB/s matches %d" % (d["matcher"][:30], d["ms"], d["haystack_GB_per_s"], d["matches"]))
PY
done

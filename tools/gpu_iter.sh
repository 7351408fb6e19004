#!/bin/bash
# one GPU iteration: full parity suite, configs timing, per-kernel launch lists of the given configs.
# usage: tools/gpu_iter.sh <tag> <configs for launch lists, e.g. "1 2 3">
TAG=${1:-it}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/${TAG}_tests.log
timeout 900 python tools/bench_configs.py --configs 1,2,3 > gpurun_out/${TAG}_configs.log 2>&1; echo "configs rc=$?"; cut -c1-330 gpurun_out/${TAG}_configs.log
for C in $2; do
  echo "== launches config $C"
  bash tools/gpu_prof_cfg.sh ${TAG}_c$C $C | awk '{v=$NF; gsub(/"/,"",v); if (v+0 > 100000) print}' | sort | uniq -c | sort -k2
done

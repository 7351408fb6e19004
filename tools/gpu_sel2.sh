#!/bin/bash
# sel2 iteration: Longest/Shortest parity tests, then per-kernel launch times of config 2.  usage: tools/gpu_sel2.sh <tag> [full]
TAG=${1:-sel2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "sel2 or literal or fuzz or early or baseline or fixture" > gpurun_out/${TAG}_tests1.log 2>&1; echo "tests1 rc=$?"
tail -5 gpurun_out/${TAG}_tests1.log
if [ "$2" == "full" ]; then
  timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"
  tail -5 gpurun_out/${TAG}_tests.log
fi
timeout 900 python tools/bench_configs.py --configs 2 > gpurun_out/${TAG}_configs.log 2>&1; echo "configs rc=$?"; cat gpurun_out/${TAG}_configs.log | cut -c1-330
bash tools/gpu_prof_cfg.sh ${TAG} 2 | awk '$NF+0 > 200000 || /top|group|tiles/' | sort | uniq -c | sort -k2 | awk '$NF+0 > 20000'

#!/bin/bash
# sel2 iteration: Longest/Shortest parity tests (fused and fused), config 2 timings for both.  usage: tools/gpu_sel2.sh <tag>
TAG=${1:-sel2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "sel2 or literal or fuzz or early or baseline or fixture" > gpurun_out/${TAG}_tests1.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/${TAG}_tests1.log
ACGPU_SEL2_FUSED=1 timeout 900 python -m pytest tests -m gpu -x -q -k "sel2 or fuzz" > gpurun_out/${TAG}_tests2.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/${TAG}_tests2.log
timeout 900 python tools/bench_configs.py --configs 2 > gpurun_out/${TAG}_configs.log 2>&1; echo "configs rc=$?"; cut -c1-250 gpurun_out/${TAG}_configs.log
ACGPU_SEL2_FUSED=1 timeout 900 python tools/bench_configs.py --configs 2 > gpurun_out/${TAG}_configs2.log 2>&1; echo "configs fused rc=$?"; cut -c1-250 gpurun_out/${TAG}_configs2.log
bash tools/gpu_prof_cfg.sh ${TAG} 2 | awk '{v=$NF; gsub(/"/,"",v); if (v+0 > 100000) print}' | sort | uniq -c | sort -k2

#!/bin/bash
# Session-6 record (short GPU budget): parity suite incl. the C++ mirror, smoke, bench (both arms).  usage: tools/gpu_s6.sh <tag> [bench]
TAG=${1:-s6}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${TAG}_tests.log; tail -4 gpurun_out/${TAG}_tests.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/${TAG}_smoke.log
if [ "$2" == "bench" ]; then
  timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench.json
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_ref.json
fi

#!/bin/bash
# sel2 bring-up: the new tests first, then the whole GPU suite, then configs 2 timing.  usage: tools/gpu_sel2.sh <tag>
TAG=${1:-sel2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "sel2 or literal or fuzz or early or baseline" > gpurun_out/${TAG}_tests1.log 2>&1; echo "tests1 rc=$?"
tail -25 gpurun_out/${TAG}_tests1.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/${TAG}_tests.log
timeout 900 python tools/bench_configs.py --configs 2 > gpurun_out/${TAG}_configs.log 2>&1; echo "configs rc=$?"; cat gpurun_out/${TAG}_configs.log | cut -c1-400

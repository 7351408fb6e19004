#!/bin/bash
# paired continuation-mask gather: tier-path parity + short device-timed bench + launch list
mkdir -p gpurun_out
TAG=${1:-r2k}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "tier or full_1m or baseline_configs or sel2 or fuzz or literal or random_dictionaries or compact or range_shards" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${TAG}_tests.log
VARIANTS="$VARIANTS" LAUNCHES=1 tools/gpu_exp.sh $TAG

#!/bin/bash
# slab runs (L2-resident hit masks): parity with small slabs, then a sweep of the slab size on configs[4]
mkdir -p gpurun_out
TAG=${1:-r2n}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
if [ -z "$NOTESTS" ]; then
ACGPU_SLAB_ROWS=4096 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "${TESTK:-tier or full_1m or baseline_configs or fuzz or literal or compact or range_shards or readable or config4 or high_positions}" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${TAG}_tests.log
fi
SHORT="python bench.py --haystacks 1 --chars 1000000000 --steps 5 --warmup 2 --no-e2e --no-cpu-baseline"
for SPEC in ${SWEEP:-0 32768 65536 131072 262144 65536:p 131072:p 65536:g3 131072:g3}; do
  R=${SPEC%%:*}; F=""
  if [[ "$SPEC" == *:* ]]; then F=${SPEC#*:}; fi
  unset ACGPU_SLAB_PERSIST ACGPU_MASK_GEN
  export ACGPU_SLAB_ROWS=$R
  if [ "$F" == "p" ]; then export ACGPU_SLAB_PERSIST=1; fi
  if [ "$F" == "g3" ]; then export ACGPU_MASK_GEN=3; fi
  timeout 300 $SHORT > gpurun_out/${TAG}_s${R}${F}_short.json 2> gpurun_out/${TAG}_s${R}${F}_short.err; rc=$?
  python -c "import sys,json; d=json.loads(open('gpurun_out/${TAG}_s${R}${F}_short.json').read()); r=d['roofline']; print('slab $SPEC rc=$rc launch_ms %.3f frac %.3f matches %d parity %s' % (r['launch_ms'], r['frac'], d['matches_per_step'], d.get('parity',{}).get('ok')))" || tail -3 gpurun_out/${TAG}_s${R}${F}_short.err
done

#!/bin/bash
# generation 5 (k_tier_fused): parity (also with tiny tickets / rings so every wait path runs), then a tuning sweep on configs[4]
mkdir -p gpurun_out
TAG=${1:-r2o}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
K="${TESTK:-tier or full_1m or baseline_configs or fuzz or literal or compact or range_shards or readable or config4 or config1 or positions_up_to}"
if [ -z "$NOTESTS" ]; then
ACGPU_FUSE=1 ACGPU_FUSE_MIN_ROWS=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "$K" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests(default tuning, every size fused) rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
ACGPU_FUSE=1 ACGPU_FUSE_MIN_ROWS=1 ACGPU_FUSE_ROWS=3 ACGPU_FUSE_RING_MB=1 ACGPU_FUSE_HIGH_MB=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "$K" > gpurun_out/${TAG}_tests_tiny.log 2>&1; echo "tests(3-row tickets, 1 MB ring) rc=$?"; tail -3 gpurun_out/${TAG}_tests_tiny.log
fi
SHORT="python bench.py --haystacks 1 --chars 1000000000 --steps 5 --warmup 2 --no-e2e --no-cpu-baseline"
for SPEC in ${SWEEP:-off 16:64:32 8:64:32 32:64:32 16:64:16 16:64:48 16:96:64 16:32:16}; do
  unset ACGPU_FUSE ACGPU_FUSE_ROWS ACGPU_FUSE_RING_MB ACGPU_FUSE_HIGH_MB
  if [ "$SPEC" == "off" ]; then export ACGPU_FUSE=0; else
    IFS=: read R RING HIGH <<< "$SPEC"; export ACGPU_FUSE=1 ACGPU_FUSE_ROWS=$R ACGPU_FUSE_RING_MB=$RING ACGPU_FUSE_HIGH_MB=$HIGH; fi
  N=${SPEC//:/_}
  timeout 200 $SHORT > gpurun_out/${TAG}_f${N}_short.json 2> gpurun_out/${TAG}_f${N}_short.err; rc=$?
  python -c "import sys,json; d=json.loads(open('gpurun_out/${TAG}_f${N}_short.json').read()); r=d['roofline']; print('fuse $SPEC rc=$rc launch_ms %.3f frac %.3f matches %d launches %s' % (r['launch_ms'], r['frac'], d['matches_per_step'], d.get('gpu_launches')))" || tail -3 gpurun_out/${TAG}_f${N}_short.err
done

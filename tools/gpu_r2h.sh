#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?"; tail -16 gpurun_out/r2h_tests.log
for V in product wc3 wc5 wc6; do
  unset ACGPU_LIB
  if [ "$V" != "product" ]; then export ACGPU_LIB=$PWD/ahocorasick_b200/variants/libacgpu_$V.so; fi
  timeout 300 python tools/bench_configs.py --configs 5 --scale 0.25 --e2e-chars 20000000 2>/dev/null | python -c "
import sys,json
for ln in sys.stdin:
    d=json.loads(ln); print('$V', d['matcher'][:34], '%.2f ms %.1f GB/s'%(d['ms'], d['haystack_GB_per_s']))"
done

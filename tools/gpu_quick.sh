#!/bin/bash
# Quick GPU check: parity tests (optional), short bench + per-kernel launch times for the product library and for any
# experiment builds (VARIANTS="w32 ..." -> ahocorasick_b200/variants/libacgpu_<v>.so), optional full ncu capture.
# usage: [VARIANTS="a b"] tools/gpu_quick.sh <tag> [tests|notests] [kernel-regex-for-full-capture] [variant-for-capture]
TAG=${1:-q}
mkdir -p gpurun_out
if [ "$2" == "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${TAG}_tests.log
  tail -5 gpurun_out/${TAG}_tests.log
fi
SHORT="python bench.py --haystacks 2 --chars 1000000000 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
NCU="python bench.py --haystacks 1 --chars 1000000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
# a variant is NAME (library ahocorasick_b200/variants/libacgpu_NAME.so) or NAME:ENV=VAL (product library, one env knob)
for SPEC in product $VARIANTS; do
  V=${SPEC%%:*}; KNOB=""
  if [[ "$SPEC" == *:* ]]; then KNOB=${SPEC#*:}; fi
  unset ACGPU_LIB
  if [ "$V" != "product" ] && [ -z "$KNOB" ]; then export ACGPU_LIB=$PWD/ahocorasick_b200/variants/libacgpu_$V.so; fi
  if [ -n "$KNOB" ]; then export "$KNOB"; fi
  echo "=== $SPEC"
  timeout 600 $SHORT > gpurun_out/${TAG}_${V}_short.json 2> gpurun_out/${TAG}_${V}_short.err; echo "short rc=$?"
  python -c "import sys,json; d=json.loads(open('gpurun_out/${TAG}_${V}_short.json').read()); r=d['roofline']; print('launch_ms %.3f frac %.3f matches %d' % (r['launch_ms'], r['frac'], d['matches_per_step']))"
  tail -3 gpurun_out/${TAG}_${V}_short.err
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_tier|k_row|k_ac' --csv --log-file gpurun_out/${TAG}_${V}_launches.csv $NCU > gpurun_out/${TAG}_${V}_launches.log 2>&1
  grep -E "k_tier|k_row|k_ac" gpurun_out/${TAG}_${V}_launches.csv | awk -F'","' '{print substr($5,1,50), $NF}' | tail -4
  if [ -n "$KNOB" ]; then unset "${KNOB%%=*}"; fi
done
if [ -n "$3" ]; then
  if [ -n "$4" ]; then export ACGPU_LIB=$PWD/ahocorasick_b200/variants/libacgpu_$4.so; else unset ACGPU_LIB; fi
  for KR in $(echo $3 | tr ',' ' '); do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KR -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_$KR $NCU > gpurun_out/${TAG}_prof_$KR.log 2>&1
    ls -la gpurun_out/${TAG}_prof_$KR.ncu-rep
  done
fi

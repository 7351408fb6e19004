#!/bin/bash
# generation-4 mask kernel (k_tier_pair) + leaner k_tier_emit: parity, A/B against generation 3, launch list, full captures
mkdir -p gpurun_out
TAG=${1:-r2l}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
if [ -z "$NOTESTS" ]; then
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "${TESTK:-tier or full_1m or baseline_configs or sel2 or fuzz or literal or random_dictionaries or compact or range_shards or readable}" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${TAG}_tests.log
fi
SHORT="python bench.py --haystacks 1 --chars 1000000000 --steps 5 --warmup 2 --no-e2e --no-cpu-baseline"
NCU="python bench.py --haystacks 1 --chars 1000000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
for SPEC in product $VARIANTS; do
  V=${SPEC%%:*}; KNOB=""
  if [[ "$SPEC" == *:* ]]; then KNOB=${SPEC#*:}; fi
  unset ACGPU_LIB
  if [ "$V" != "product" ] && [ -z "$KNOB" ]; then export ACGPU_LIB=$PWD/ahocorasick_b200/variants/libacgpu_$V.so; fi
  if [ -n "$KNOB" ]; then export "$KNOB"; fi
  timeout 300 $SHORT > gpurun_out/${TAG}_${V}_short.json 2> gpurun_out/${TAG}_${V}_short.err; rc=$?
  python -c "import sys,json; d=json.loads(open('gpurun_out/${TAG}_${V}_short.json').read()); r=d['roofline']; print('$SPEC rc=$rc launch_ms %.3f frac %.3f matches %d' % (r['launch_ms'], r['frac'], d['matches_per_step']))" || tail -3 gpurun_out/${TAG}_${V}_short.err
  if [ -n "$LAUNCHES" ]; then
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_${V}_launches.csv $NCU > gpurun_out/${TAG}_${V}_launches.log 2>&1
    grep -E '"k_|"void' gpurun_out/${TAG}_${V}_launches.csv | grep -v "at::" | awk -F'","' '{print substr($5,1,60), $NF}' | tail -3
  fi
  if [ -n "$KNOB" ]; then unset "${KNOB%%=*}"; fi
done
unset ACGPU_LIB
for KR in $FULL; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KR -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_$KR $NCU > gpurun_out/${TAG}_prof_$KR.log 2>&1
  ncu -i gpurun_out/${TAG}_prof_$KR.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_${KR}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_prof_$KR.ncu-rep --page source --csv > gpurun_out/${TAG}_prof_${KR}_source.csv 2>/dev/null
  ls -la gpurun_out/${TAG}_prof_$KR.ncu-rep
done

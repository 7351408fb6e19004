#!/bin/bash
# Kernel experiment call: short device-timed bench of the product library and of experiment builds
# (VARIANTS="a b" -> ahocorasick_b200/variants/libacgpu_<v>.so), then optional ncu --set full captures (raw CSV kept).
# usage: [VARIANTS="a b"] [FULL="k_tier_mask k_tier_emit"] [FULLVAR=name] tools/gpu_exp.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
SHORT="python bench.py --haystacks 1 --chars 1000000000 --steps 5 --warmup 2 --no-e2e --no-cpu-baseline"
NCU="python bench.py --haystacks 1 --chars 1000000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
for V in product $VARIANTS; do
  unset ACGPU_LIB
  if [ "$V" != "product" ]; then export ACGPU_LIB=$PWD/ahocorasick_b200/variants/libacgpu_$V.so; fi
  timeout 300 $SHORT > gpurun_out/${TAG}_${V}_short.json 2> gpurun_out/${TAG}_${V}_short.err; rc=$?
  python -c "import sys,json; d=json.loads(open('gpurun_out/${TAG}_${V}_short.json').read()); r=d['roofline']; print('$V rc=$rc launch_ms %.3f frac %.3f matches %d clk %s' % (r['launch_ms'], r['frac'], d['matches_per_step'], d['clocks']['sm_mhz']))" || tail -3 gpurun_out/${TAG}_${V}_short.err
done
unset ACGPU_LIB
if [ -n "$FULLVAR" ]; then export ACGPU_LIB=$PWD/ahocorasick_b200/variants/libacgpu_$FULLVAR.so; fi
for KR in $FULL; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KR -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_$KR $NCU > gpurun_out/${TAG}_prof_$KR.log 2>&1
  ncu -i gpurun_out/${TAG}_prof_$KR.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_${KR}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_prof_$KR.ncu-rep --page source --csv > gpurun_out/${TAG}_prof_${KR}_source.csv 2>/dev/null
  ls -la gpurun_out/${TAG}_prof_$KR.ncu-rep
  if [ $(stat -c %s gpurun_out/${TAG}_prof_$KR.ncu-rep) -gt 20000000 ]; then rm gpurun_out/${TAG}_prof_$KR.ncu-rep; fi
done
if [ -n "$LAUNCHES" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --haystacks 1 --chars 1000000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
  grep -E '"k_|"void' gpurun_out/${TAG}_launches.csv | awk -F'","' '{print substr($5,1,60), $NF}' | tail -8
fi

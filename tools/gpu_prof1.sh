#!/bin/bash
# one ncu --set full capture of a kernel of the configs[4] bench: tools/gpu_prof1.sh <tag> <kernel regex>
mkdir -p gpurun_out
TAG=$1; KR=$2
NCU="python bench.py --haystacks 1 --chars 1000000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KR -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_$KR $NCU > gpurun_out/${TAG}_prof_$KR.log 2>&1
ncu -i gpurun_out/${TAG}_prof_$KR.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_${KR}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_prof_$KR.ncu-rep --page source --csv > gpurun_out/${TAG}_prof_${KR}_source.csv 2>/dev/null
ls -la gpurun_out/${TAG}_prof_$KR.ncu-rep
if [ $(stat -c %s gpurun_out/${TAG}_prof_$KR.ncu-rep) -gt 30000000 ]; then rm gpurun_out/${TAG}_prof_$KR.ncu-rep; fi
tail -3 gpurun_out/${TAG}_prof_$KR.log

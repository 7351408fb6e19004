#!/bin/bash
# full captures of the configs[2] kernels (Longest: mirrored k_tier_pair, k_sel2_map / emit / values) for the record
mkdir -p gpurun_out
TAG=${1:-r5m}
CMD="python tools/bench_configs.py --configs 2 --scale 0.25 --steps 1 --warmup 1 --e2e-chars 1000000"
for KR in k_tier_pair k_sel2_map k_sel2_emit k_sel2_values; do
  timeout 300 ncu --set full --clock-control none -k regex:$KR -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_$KR $CMD > gpurun_out/${TAG}_prof_$KR.log 2>&1
  ncu -i gpurun_out/${TAG}_prof_$KR.ncu-rep --page details > gpurun_out/${TAG}_ncu_full_$KR.txt 2>/dev/null
  rm -f gpurun_out/${TAG}_prof_$KR.ncu-rep
  grep -E "^\s+(Duration|Executed Ipc Active|Issue Slots Busy|DRAM Throughput|L2 Hit Rate|Registers Per Thread|Achieved Occupancy)" gpurun_out/${TAG}_ncu_full_$KR.txt | tr -s ' ' | tr '\n' ';'; echo " <- $KR"
done

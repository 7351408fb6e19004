#!/bin/bash
# per-kernel launch times (+ optional full captures) of one bench_configs config.  usage: tools/gpu_prof_cfg.sh <tag> <config> [kernel-regex,...]
TAG=${1:-p}; CFG=${2:-2}
mkdir -p gpurun_out
CMD="python tools/bench_configs.py --configs $CFG --scale 0.5 --steps 1 --warmup 1 --e2e-chars 1000000"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_tier|k_sel|k_row|k_fwd|k_ac|k_ww' --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
grep -E "k_tier|k_sel|k_row|k_fwd|k_ac|k_ww" gpurun_out/${TAG}_launches.csv | awk -F'","' '{print substr($5,1,60), $NF}' | tail -40
if [ -n "$3" ]; then
  for KR in $(echo $3 | tr ',' ' '); do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KR -s ${SKIP:-0} -c 1 -f -o gpurun_out/${TAG}_prof_$KR $CMD > gpurun_out/${TAG}_prof_$KR.log 2>&1
    ls -la gpurun_out/${TAG}_prof_$KR.ncu-rep
  done
fi

#!/bin/bash
# fused kernel: sweep with DRAM bytes per variant (ncu, 3 metrics) next to the event-timed bench
mkdir -p gpurun_out
TAG=${1:-r2r}
SHORT="python bench.py --haystacks 1 --chars 1000000000 --steps 5 --warmup 2 --no-e2e --no-cpu-baseline"
NCU="python bench.py --haystacks 1 --chars 1000000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
for SPEC in ${SWEEP}; do
  unset ACGPU_FUSE ACGPU_FUSE_ROWS ACGPU_FUSE_RING_MB ACGPU_FUSE_HIGH_MB
  if [ "$SPEC" == "off" ]; then export ACGPU_FUSE=0; else
    IFS=: read R RING HIGH <<< "$SPEC"; export ACGPU_FUSE=1 ACGPU_FUSE_ROWS=$R ACGPU_FUSE_RING_MB=$RING ACGPU_FUSE_HIGH_MB=$HIGH; fi
  N=${SPEC//:/_}
  timeout 200 $SHORT > gpurun_out/${TAG}_f${N}_short.json 2> gpurun_out/${TAG}_f${N}_short.err; rc=$?
  python -c "import sys,json; d=json.loads(open('gpurun_out/${TAG}_f${N}_short.json').read()); r=d['roofline']; print('fuse $SPEC rc=$rc launch_ms %.3f frac %.3f matches %d' % (r['launch_ms'], r['frac'], d['matches_per_step']))" || tail -3 gpurun_out/${TAG}_f${N}_short.err
  if [ -n "$DRAM" ]; then
    timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_tier -s 3 -c 3 --csv --log-file gpurun_out/${TAG}_f${N}_dram.csv $NCU > /dev/null 2>&1
    grep -E '"k_tier|"void' gpurun_out/${TAG}_f${N}_dram.csv | awk -F'","' '{print "   ", substr($5,1,40), $(NF-2), $(NF-1), $NF}'
  fi
done

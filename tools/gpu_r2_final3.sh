#!/bin/bash
# closing validation of the final binary: whole GPU suite, smoke, bench both arms, configs, launch list + traffic captures
TAG=${1:-r2z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; cut -c1-160 gpurun_out/${TAG}_bench_ref.json
timeout 900 python tools/bench_configs.py --configs 0,1,2,3,5 > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err; echo "configs rc=$?"
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_configs.jsonl"):
    d = json.loads(ln)
    print("cfg %d %-36s %8.2f ms %7.1f GB/s frac %.3f e2e %5.1f GB/s stream %s" % (d["config"], d["matcher"][:36], d["ms"], d["haystack_GB_per_s"], d["roofline"]["frac"], d["e2e_GB_per_s"], d.get("readable_stream_GB_per_s")))
PY
timeout 400 python tools/bench_stream_sweep.py > gpurun_out/${TAG}_stream_sweep.jsonl 2> gpurun_out/${TAG}_stream_sweep.err; echo "sweep rc=$?"; cut -c1-330 gpurun_out/${TAG}_stream_sweep.jsonl
NCU="python bench.py --haystacks 1 --chars 1000000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_tier|k_row' --csv --log-file gpurun_out/${TAG}_launches_config4.csv $NCU > gpurun_out/${TAG}_launches.log 2>&1
grep -E "k_tier|k_row" gpurun_out/${TAG}_launches_config4.csv | awk -F'","' '{print substr($5,1,50), $NF}' | tail -3
for KR in k_tier_mask k_tier_emit; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KR -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_$KR $NCU > gpurun_out/${TAG}_prof_$KR.log 2>&1
  ncu -i gpurun_out/${TAG}_prof_$KR.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_${KR}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_prof_$KR.ncu-rep --page details > gpurun_out/${TAG}_ncu_full_$KR.txt 2>/dev/null
  rm -f gpurun_out/${TAG}_prof_$KR.ncu-rep
done
ls gpurun_out | grep ${TAG} | wc -l

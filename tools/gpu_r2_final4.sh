#!/bin/bash
# last validation of the final binary: whole GPU suite, smoke, short bench, config 3, WholeWord captures
TAG=${1:-r2zz}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=3 > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench.json').read()); print(d['value'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['parity_checked'], d['clocks'])"
timeout 600 python tools/bench_configs.py --configs 3 > gpurun_out/${TAG}_cfg3.jsonl 2> gpurun_out/${TAG}_cfg3.err
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_cfg3.jsonl"):
    d = json.loads(ln)
    print("cfg %d %-36s %8.2f ms %7.1f GB/s frac %.3f e2e %5.1f GB/s stream %s launches %s" % (d["config"], d["matcher"][:36], d["ms"], d["haystack_GB_per_s"], d["roofline"]["frac"], d["e2e_GB_per_s"], d.get("readable_stream_GB_per_s"), d["launches_per_match"]))
PY
CFG3="python tools/bench_configs.py --configs 3 --scale 0.5 --steps 1 --warmup 1 --e2e-chars 1000000"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_ww3|k_row' -c 40 --csv --log-file gpurun_out/${TAG}_launches_config3.csv $CFG3 > /dev/null 2>&1
grep -E "k_ww3|k_row" gpurun_out/${TAG}_launches_config3.csv | awk -F'","' '{gsub(/"/,"",$NF); if ($NF+0 > 15000) print substr($5,1,40), $NF}' | head -8
for KR in k_ww3_hits k_ww3_emit; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KR -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_$KR $CFG3 > gpurun_out/${TAG}_prof_$KR.log 2>&1
  ncu -i gpurun_out/${TAG}_prof_$KR.ncu-rep --page details > gpurun_out/${TAG}_ncu_full_$KR.txt 2>/dev/null
  rm -f gpurun_out/${TAG}_prof_$KR.ncu-rep
done

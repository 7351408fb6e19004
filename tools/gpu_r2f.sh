#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "wide or fuzz_small or random_dictionaries or literal or readable" > gpurun_out/r2f_tests.log 2>&1; tail -30 gpurun_out/r2f_tests.log
timeout 600 python tools/bench_configs.py --configs 5 --scale 0.5 > gpurun_out/r2f_cfg5.jsonl 2> gpurun_out/r2f_cfg5.err; tail -3 gpurun_out/r2f_cfg5.err
python - <<'PY'
import json
for ln in open("gpurun_out/r2f_cfg5.jsonl"):
    d = json.loads(ln)
    print("cfg %d %-50s %8.2f ms %7.1f GB/s frac %.3f e2e %5.1f GB/s stream %s launches %d" % (d["config"], d["matcher"], d["ms"], d["haystack_GB_per_s"], d["roofline"]["frac"], d["e2e_GB_per_s"], d.get("readable_stream_GB_per_s"), d["launches_per_match"]))
PY
ACGPU_FORCE_GEN1=1 timeout 600 python tools/bench_configs.py --configs 5 --scale 0.05 2>/dev/null | python -c "
import sys,json
for ln in sys.stdin:
    d=json.loads(ln); print('gen1', d['matcher'][:30], '%.1f GB/s'%d['haystack_GB_per_s'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_cfg5.csv python tools/bench_configs.py --configs 5 --scale 0.25 --steps 1 --warmup 1 > /dev/null 2>&1
grep -E "k_wide|k_row" gpurun_out/r2f_launches_cfg5.csv | awk -F'","' '{print substr($5,1,60), $NF}' | tail -6
bash tools/gpu_exp.sh r2f

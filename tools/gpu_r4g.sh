#!/bin/bash
# quick: WholeWord parity subset + config 3 timing + instruction count of k_ww3_hits
mkdir -p gpurun_out
TAG=${1:-r4g}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "wholeword or ww or readable or baseline_configs or word" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${TAG}_tests.log
timeout 600 python tools/bench_configs.py --configs 3 --scale 0.5 --e2e-chars 1000000 > gpurun_out/${TAG}_cfg.jsonl 2> gpurun_out/${TAG}_cfg.err
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_cfg.jsonl"):
    d = json.loads(ln)
    print("cfg %d %-36s %8.3f ms %7.1f GB/s frac %.3f matches %d" % (d["config"], d["matcher"][:36], d["ms"], d["haystack_GB_per_s"], d["roofline"]["frac"], d["matches"]))
PY
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_lsu.sum --clock-control none -k regex:'k_ww3_hits' -c 2 --csv python tools/bench_configs.py --configs 3 --scale 0.5 --steps 1 --warmup 1 --e2e-chars 1000000 2>/dev/null | grep k_ww3 | awk -F'","' '{print $(NF-3), $NF}'

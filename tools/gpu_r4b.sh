#!/bin/bash
# WholeWord generation 3: parity subset, config 3 timing, per-kernel launch list, optional full capture of k_ww3_hits
mkdir -p gpurun_out
TAG=${1:-r4b}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "wholeword or ww or readable or config3 or baseline_configs or word" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
timeout 600 python tools/bench_configs.py --configs 3 --scale 0.5 --e2e-chars 1000000 > gpurun_out/${TAG}_cfg.jsonl 2> gpurun_out/${TAG}_cfg.err
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_cfg.jsonl"):
    d = json.loads(ln)
    print("cfg %d %-36s %8.3f ms %7.1f GB/s frac %.3f matches %d" % (d["config"], d["matcher"][:36], d["ms"], d["haystack_GB_per_s"], d["roofline"]["frac"], d["matches"]))
PY
tools/gpu_prof_cfg.sh ${TAG} 3 $2 | grep -v wwl_starts | grep -v k_sel | tail -14

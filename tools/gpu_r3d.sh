#!/bin/bash
# Map value look-ups with four gathers in flight: parity (Maps) + configs 1, 2
mkdir -p gpurun_out
TAG=${1:-r3d}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "map or Map or sel2 or baseline_configs or config1 or config2 or readable or fuzz" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
timeout 600 python tools/bench_configs.py --configs 1,2 --scale 0.5 --e2e-chars 1000000 > gpurun_out/${TAG}_cfg.jsonl 2> gpurun_out/${TAG}_cfg.err
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_cfg.jsonl"):
    d = json.loads(ln)
    print("cfg %d %-28s %8.3f ms %7.1f GB/s frac %.3f" % (d["config"], d["matcher"][:28], d["ms"], d["haystack_GB_per_s"], d["roofline"]["frac"]))
PY

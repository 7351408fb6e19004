#!/bin/bash
# k_ww3 per-hit rework (start bitmap, value scratch, count kernel): parity + sparse (configs[3]) + dense (every word a keyword)
mkdir -p gpurun_out
TAG=${1:-r5j}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "wholeword or ww or readable or config3 or baseline_configs or word or shard" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
timeout 600 python tools/bench_configs.py --configs 3 --scale 0.5 --steps 5 --e2e-chars 1000000 > gpurun_out/${TAG}_cfg3.jsonl 2> gpurun_out/${TAG}_cfg3.err
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_cfg3.jsonl"):
    d = json.loads(ln)
    print("cfg %d %-36s %10d chars %8.3f ms %7.1f GB/s frac %.3f" % (d["config"], d["matcher"][:36], d["chars"], d["ms"], d["haystack_GB_per_s"], d["roofline"]["frac"]))
PY
timeout 300 python tools/bench_ww_dense.py > gpurun_out/${TAG}_dense.jsonl 2> gpurun_out/${TAG}_dense.err; cat gpurun_out/${TAG}_dense.jsonl

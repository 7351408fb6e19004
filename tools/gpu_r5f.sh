#!/bin/bash
# child signatures in the edge info words (trie_step_sig): parity of everything that walks the hashed trie + timings of the walks
mkdir -p gpurun_out
TAG=${1:-r5f}
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "not (config4 or full_1m or positions_up_to or host_call)" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
timeout 600 python tools/bench_configs.py --configs 3 --scale 0.5 --steps 5 --e2e-chars 1000000 > gpurun_out/${TAG}_cfg3.jsonl 2> gpurun_out/${TAG}_cfg3.err
ACGPU_FORCE_GEN1=1 timeout 600 python tools/bench_configs.py --configs 2 --scale 0.05 --steps 3 --e2e-chars 1000000 > gpurun_out/${TAG}_cfg2_gen1.jsonl 2> gpurun_out/${TAG}_cfg2_gen1.err
python - <<PY
import json
for f in ("cfg3", "cfg2_gen1"):
    for ln in open("gpurun_out/${TAG}_%s.jsonl" % f):
        d = json.loads(ln)
        print("%-10s cfg %d %-36s %10d chars %8.3f ms %7.1f GB/s frac %.3f" % (f, d["config"], d["matcher"][:36], d["chars"], d["ms"], d["haystack_GB_per_s"], d["roofline"]["frac"]))
PY

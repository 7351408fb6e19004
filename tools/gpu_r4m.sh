#!/bin/bash
# configs 1-3 timing + parity subset after the classification change
mkdir -p gpurun_out
TAG=${1:-r4m}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "${KEXPR:-config1 or baseline_configs or fuzz or unicode or case or readable or ww or word}" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${TAG}_tests.log
timeout 600 python tools/bench_configs.py --configs ${CFGS:-1,2,3} --scale 0.5 --e2e-chars 1000000 > gpurun_out/${TAG}_cfg.jsonl 2> gpurun_out/${TAG}_cfg.err
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_cfg.jsonl"):
    d = json.loads(ln)
    print("cfg %d %-36s %8.3f ms %7.1f GB/s frac %.3f matches %d" % (d["config"], d["matcher"][:36], d["ms"], d["haystack_GB_per_s"], d["roofline"]["frac"], d["matches"]))
PY

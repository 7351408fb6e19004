#!/bin/bash
# wide Map values by keyword hash: parity + A/B
mkdir -p gpurun_out
TAG=${1:-r3c}
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "wide or fuzz_small or random_dictionaries or literal or unicode or full_node" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
for V in hash walk; do
  ACGPU_WIDE_VALUES=$V timeout 600 python tools/bench_configs.py --configs 5 --scale 0.5 > gpurun_out/${TAG}_cfg5_$V.jsonl 2> gpurun_out/${TAG}_cfg5_$V.err; tail -2 gpurun_out/${TAG}_cfg5_$V.err
  python - <<PY
import json
for ln in open("gpurun_out/${TAG}_cfg5_$V.jsonl"):
    d = json.loads(ln)
    print("values=$V cfg %d %-50s %8.2f ms %7.1f GB/s frac %.3f e2e %5.1f GB/s matches %d" % (d["config"], d["matcher"], d["ms"], d["haystack_GB_per_s"], d["roofline"]["frac"], d["e2e_GB_per_s"], d["matches"]))
PY
done

#!/bin/bash
# WholeWord generation 3 (kernel_ww3.cuh) against generation 2 (k_ww_scan): parity (WholeWord paths) + config 3
mkdir -p gpurun_out
TAG=${1:-r4a}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "wholeword or ww or readable or config3 or baseline_configs or literal or fuzz or shard or word" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${TAG}_tests.log
for G in 3 2; do
  export ACGPU_WW_GEN=$G
  timeout 600 python tools/bench_configs.py --configs 3 --scale 0.5 --e2e-chars 1000000 > gpurun_out/${TAG}_cfg_gen$G.jsonl 2> gpurun_out/${TAG}_cfg_gen$G.err
  python - <<PY
import json
for ln in open("gpurun_out/${TAG}_cfg_gen$G.jsonl"):
    d = json.loads(ln)
    print("gen$G cfg %d %-36s %8.3f ms %7.1f GB/s frac %.3f matches %d" % (d["config"], d["matcher"][:36], d["ms"], d["haystack_GB_per_s"], d["roofline"]["frac"], d["matches"]))
PY
  tail -2 gpurun_out/${TAG}_cfg_gen$G.err
done

#!/bin/bash
# k_ww3_hits small tickets: WholeWord parity + config 3 at three sizes + Readable sweep of config 3
mkdir -p gpurun_out
TAG=${1:-r5e}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "wholeword or ww or readable or config3 or baseline_configs or word" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
for SC in 0.001 0.01 0.5; do
  timeout 600 python tools/bench_configs.py --configs 3 --scale $SC --steps 20 --warmup 5 --e2e-chars 100000000 > gpurun_out/${TAG}_cfg_$SC.jsonl 2> gpurun_out/${TAG}_cfg_$SC.err
  python - <<PY
import json
for ln in open("gpurun_out/${TAG}_cfg_$SC.jsonl"):
    d = json.loads(ln)
    if "Longest" in d["matcher"]: continue
    print("cfg %d %-22s %10d chars %8.4f ms %7.1f GB/s frac %.3f e2e %5.1f" % (d["config"], d["matcher"][:22], d["chars"], d["ms"], d["haystack_GB_per_s"], d["roofline"]["frac"], d["e2e_GB_per_s"]))
PY
done
timeout 300 python tools/bench_stream_sweep.py --configs 3 > gpurun_out/${TAG}_sweep.jsonl 2>/dev/null; cut -c1-330 gpurun_out/${TAG}_sweep.jsonl

#!/bin/bash
# HostCall::run_chain (Longest / Shortest host calls as pipelined chain shards): parity + e2e of config 2 against upload-scan-download
mkdir -p gpurun_out
TAG=${1:-r5c}
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -x -q -k "host_call_pipelines or config2 or chain_shards or large_haystack" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
for MODE in 1 0; do
  ACGPU_HOST_CHAIN=$MODE timeout 600 python tools/bench_configs.py --configs 2 --scale 0.25 --steps 2 --e2e-chars 200000000 > gpurun_out/${TAG}_cfg_chain$MODE.jsonl 2> gpurun_out/${TAG}_cfg_chain$MODE.err
  python - <<PY
import json
for ln in open("gpurun_out/${TAG}_cfg_chain$MODE.jsonl"):
    d = json.loads(ln)
    print("host_chain=$MODE cfg %d %-20s e2e %6.1f GB/s on %d chars (d2h %d bytes)" % (d["config"], d["matcher"][:20], d["e2e_GB_per_s"], d["e2e_chars"], d["e2e_d2h_bytes"]))
PY
done

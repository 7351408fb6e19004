#!/bin/bash
# Readable sweeps after "free on the allocating stream": every family, three times
mkdir -p gpurun_out
TAG=${1:-r5h}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "readable or stream" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${TAG}_tests.log
for i in 1 2 3; do
  timeout 600 python tools/bench_stream_sweep.py > gpurun_out/${TAG}_sweep_$i.jsonl 2> /dev/null
  python - <<PY
import json
for ln in open("gpurun_out/${TAG}_sweep_$i.jsonl"):
    d = json.loads(ln); g = d["GB_per_s"]
    print("run $i %-20s" % d["matcher"][:20], " ".join("%s=%.1f" % (k.replace("adaptive 2^16..2^24", "adp"), v) for k, v in g.items()))
PY
done

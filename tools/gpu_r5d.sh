#!/bin/bash
# small tickets for short haystacks (mask_chunk_rows): tier-path parity + config 0 / stream sweep with and without
mkdir -p gpurun_out
TAG=${1:-r5d}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "tier_path or baseline_configs or literal_cases or fuzz or config1 or config2_full or sel2 or chain or compact or fixture" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
for ROWS in 0 32; do
  ACGPU_CHUNK_ROWS=$ROWS timeout 600 python tools/bench_configs.py --configs 0,1,2 --scale 0.1 --steps 20 --warmup 5 --e2e-chars 100000000 > gpurun_out/${TAG}_cfg_rows$ROWS.jsonl 2> gpurun_out/${TAG}_cfg_rows$ROWS.err
  python - <<PY
import json
for ln in open("gpurun_out/${TAG}_cfg_rows$ROWS.jsonl"):
    d = json.loads(ln)
    print("chunk_rows=$ROWS cfg %d %-22s %9d chars %8.4f ms %7.1f GB/s frac %.3f e2e %5.1f" % (d["config"], d["matcher"][:22], d["chars"], d["ms"], d["haystack_GB_per_s"], d["roofline"]["frac"], d["e2e_GB_per_s"]))
PY
done

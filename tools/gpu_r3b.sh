#!/bin/bash
# generation 3 (k_tier_mask) against generation 4 (k_tier_pair) on configs 1, 2 and 4
mkdir -p gpurun_out
TAG=${1:-r3b}
for G in 4 3; do
  export ACGPU_MASK_GEN=$G
  timeout 600 python tools/bench_configs.py --configs 1,2 --scale 0.5 --e2e-chars 1000000 > gpurun_out/${TAG}_cfg_gen$G.jsonl 2> gpurun_out/${TAG}_cfg_gen$G.err
  python - <<PY
import json
for ln in open("gpurun_out/${TAG}_cfg_gen$G.jsonl"):
    d = json.loads(ln)
    print("gen$G cfg %d %-28s %8.3f ms %7.1f GB/s frac %.3f" % (d["config"], d["matcher"][:28], d["ms"], d["haystack_GB_per_s"], d["roofline"]["frac"]))
PY
  for rep in 1 2; do
  timeout 300 python bench.py --haystacks 1 --chars 1000000000 --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_c4_gen${G}_$rep.json 2>/dev/null
  python -c "import json; d=json.loads(open('gpurun_out/${TAG}_c4_gen${G}_$rep.json').read()); r=d['roofline']; print('gen$G cfg 4 run $rep launch_ms %.3f frac %.3f' % (r['launch_ms'], r['frac']))"
  done
done

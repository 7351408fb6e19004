#!/bin/bash
# k_ww3_hits experiment builds (VARIANTS="swz chunk64 sts" -> ahocorasick_b200/variants/libacgpu_<v>.so) on config 3
mkdir -p gpurun_out
TAG=${1:-r4k}
for V in product $VARIANTS; do
  unset ACGPU_LIB
  if [ "$V" != "product" ]; then export ACGPU_LIB=$PWD/ahocorasick_b200/variants/libacgpu_$V.so; fi
  timeout 300 python tools/bench_configs.py --configs 3 --scale 0.5 --steps 8 --e2e-chars 1000000 > gpurun_out/${TAG}_${V}.jsonl 2> gpurun_out/${TAG}_${V}.err
  python - <<PY
import json
for ln in open("gpurun_out/${TAG}_${V}.jsonl"):
    d = json.loads(ln)
    if "Longest" in d["matcher"]: continue
    print("$V %-22s %8.3f ms %7.1f GB/s frac %.3f matches %d" % (d["matcher"][:22], d["ms"], d["haystack_GB_per_s"], d["roofline"]["frac"], d["matches"]))
PY
done

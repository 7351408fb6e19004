#!/bin/bash
# wide path generation 2 (k_wide_tile): parity, A/B against k_wide_mask on the English-like dictionary, launch list, one full capture
mkdir -p gpurun_out
TAG=${1:-r2t}
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "wide or fuzz_small or random_dictionaries or literal" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${TAG}_tests.log
for G in 2 2notail 1; do
  unset ACGPU_WIDE_TAIL; if [ "$G" == "2notail" ]; then export ACGPU_WIDE_TAIL=0; fi
  ACGPU_WIDE_GEN=${G:0:1} timeout 600 python tools/bench_configs.py --configs 5 --scale 0.5 > gpurun_out/${TAG}_cfg5_gen$G.jsonl 2> gpurun_out/${TAG}_cfg5_gen$G.err; tail -3 gpurun_out/${TAG}_cfg5_gen$G.err
  python - <<PY
import json
for ln in open("gpurun_out/${TAG}_cfg5_gen$G.jsonl"):
    d = json.loads(ln)
    print("gen$G cfg %d %-50s %8.2f ms %7.1f GB/s frac %.3f e2e %5.1f GB/s matches %d" % (d["config"], d["matcher"], d["ms"], d["haystack_GB_per_s"], d["roofline"]["frac"], d["e2e_GB_per_s"], d["matches"]))
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/${TAG}_launches_cfg5.csv python tools/bench_configs.py --configs 5 --scale 0.25 --steps 1 --warmup 1 > /dev/null 2>&1
grep -E "k_wide|k_row" gpurun_out/${TAG}_launches_cfg5.csv | awk -F'","' '{print substr($5,1,60), $NF}' | head -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wide_tile -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_k_wide_tile python tools/bench_configs.py --configs 5 --scale 0.25 --steps 1 --warmup 1 > gpurun_out/${TAG}_prof.log 2>&1
ncu -i gpurun_out/${TAG}_prof_k_wide_tile.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_k_wide_tile_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_prof_k_wide_tile_raw.csv | grep -v "stalled_\(drain\|membar\|misc\|lg_thr\|sleep\)" | head -50

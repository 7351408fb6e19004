#!/bin/bash
# One-call session record on a B200 box (about 9 GPU-minutes): parity suite, smoke, bench (both arms), every BASELINE config,
# Readable block-size sweep, ncu launch list of the bench command.  usage: gpurun --timeout 900 -- 'bash tools/gpu_record.sh <tag>'
# Copy what should be judged from gpurun_out/ into profiles/ afterwards (see profiles/r01_s6_summary.md for the layout).
TAG=${1:-rec}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 600 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_tests.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_ref.json
timeout 600 python tools/bench_configs.py > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err; echo "configs rc=$?"
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_configs.jsonl"):
    d = json.loads(ln)
    print("cfg %d %-36s %8.2f ms %7.1f GB/s frac %.3f e2e %5.1f GB/s stream %s" % (d["config"], d["matcher"], d["ms"], d["haystack_GB_per_s"], d["roofline"]["frac"], d["e2e_GB_per_s"], d.get("readable_stream_GB_per_s")))
PY
timeout 300 python tools/bench_stream_sweep.py > gpurun_out/${TAG}_stream_sweep.jsonl 2> gpurun_out/${TAG}_stream_sweep.err; echo "sweep rc=$?"; cat gpurun_out/${TAG}_stream_sweep.jsonl
NCU="python bench.py --haystacks 1 --chars 1000000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_tier|k_row' --csv --log-file gpurun_out/${TAG}_launches.csv $NCU > gpurun_out/${TAG}_launches.log 2>&1
grep -E "k_tier|k_row" gpurun_out/${TAG}_launches.csv | awk -F'","' '{print substr($5,1,50), $NF}' | tail -3

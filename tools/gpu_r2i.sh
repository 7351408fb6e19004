#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cpp_host.py -x -q -k "chain_shards or readable or stream or cpp or committed_fixture or literal" > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2i_tests.log
timeout 300 python tools/bench_stream_sweep.py > gpurun_out/r2i_stream_sweep.jsonl 2> gpurun_out/r2i_stream_sweep.err; echo "sweep rc=$?"; cat gpurun_out/r2i_stream_sweep.jsonl | cut -c1-600
ACGPU_STREAM_SYNC=1 timeout 300 python tools/bench_stream_sweep.py > gpurun_out/r2i_stream_sweep_sync.jsonl 2>/dev/null; cat gpurun_out/r2i_stream_sweep_sync.jsonl | cut -c1-600
for c in 0 1 2 3; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 >> gpurun_out/r2i_bench_configs.jsonl 2>> gpurun_out/r2i_bench_configs.err; done
python - <<'PY'
import json
for ln in open("gpurun_out/r2i_bench_configs.jsonl"):
    d = json.loads(ln)
    print("%-42s %8.2f ms %7.1f GB/s frac %.3f e2e %5.1f GB/s stream %s cpu %s" % (d["config"]["matcher"], d["ms_per_step"], d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"].get("readable_stream_GB_per_s"), (d["cpu_baseline"] or {}).get("value")))
PY
tail -3 gpurun_out/r2i_bench_configs.err

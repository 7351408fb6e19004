#!/bin/bash
# chain-block streaming (Longest / Shortest Readable feeds on the start-mask path): parity + stream sweep of config 2, both generations
mkdir -p gpurun_out
TAG=${1:-r4n}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_cpp_host.py -x -q -k "readable or stream or fixture or early_stop" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
for G in 2 1; do
  export ACGPU_STREAM_CHAIN_GEN=$G
  timeout 600 python tools/bench_stream_sweep.py --configs 2 > gpurun_out/${TAG}_sweep_gen$G.jsonl 2> gpurun_out/${TAG}_sweep_gen$G.err || tail -3 gpurun_out/${TAG}_sweep_gen$G.err
  echo "gen $G"; cut -c1-400 gpurun_out/${TAG}_sweep_gen$G.jsonl | tail -8
done

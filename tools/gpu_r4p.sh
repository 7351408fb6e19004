#!/bin/bash
# pipelined chain-shard feeds: parity (Readable / stream tests) + stream sweep of every family, sync vs pipelined chain feeds
mkdir -p gpurun_out
TAG=${1:-r4p}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_cpp_host.py -x -q -k "readable or stream or fixture or early_stop" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
timeout 600 python tools/bench_stream_sweep.py > gpurun_out/${TAG}_sweep.jsonl 2> gpurun_out/${TAG}_sweep.err || tail -3 gpurun_out/${TAG}_sweep.err
cut -c1-330 gpurun_out/${TAG}_sweep.jsonl
ACGPU_STREAM_SYNC=1 timeout 600 python tools/bench_stream_sweep.py --configs 2 > gpurun_out/${TAG}_sweep_sync.jsonl 2> gpurun_out/${TAG}_sweep_sync.err
echo "sync feeds:"; cut -c1-330 gpurun_out/${TAG}_sweep_sync.jsonl

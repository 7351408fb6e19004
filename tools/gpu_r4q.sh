#!/bin/bash
# chain-shard feeds, pipelined against synchronous, alternating (run-to-run noise of the host side)
mkdir -p gpurun_out
TAG=${1:-r4q}
for i in 1 2; do
  for MODE in 0 1; do
    ACGPU_STREAM_SYNC=$MODE timeout 600 python tools/bench_stream_sweep.py --configs 2 > gpurun_out/${TAG}_sweep_sync${MODE}_$i.jsonl 2> /dev/null
    echo "sync=$MODE run $i: $(cut -c60-330 gpurun_out/${TAG}_sweep_sync${MODE}_$i.jsonl)"
  done
done

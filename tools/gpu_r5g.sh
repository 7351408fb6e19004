#!/bin/bash
# Readable feeds of config 1 (AhoCorasickMap), pipelined against synchronous, alternating: is the pipelined path what is slow on some boxes?
mkdir -p gpurun_out
TAG=${1:-r5g}
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv,noheader; nproc; cat /proc/loadavg
for i in 1 2; do
  for MODE in 0 1; do
    ACGPU_STREAM_SYNC=$MODE timeout 600 python tools/bench_stream_sweep.py --configs 1 > gpurun_out/${TAG}_sweep_sync${MODE}_$i.jsonl 2> /dev/null
    echo "sync=$MODE run $i: $(cut -c60-330 gpurun_out/${TAG}_sweep_sync${MODE}_$i.jsonl)"
  done
done
cat /proc/loadavg

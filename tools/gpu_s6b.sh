#!/bin/bash
# Session-6 second record: full parity suite with durations, Readable block-size sweep.  usage: tools/gpu_s6b.sh <tag>
TAG=${1:-s6b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${TAG}_tests.log; tail -22 gpurun_out/${TAG}_tests.log
timeout 400 python tools/bench_stream_sweep.py > gpurun_out/${TAG}_stream_sweep.jsonl 2> gpurun_out/${TAG}_stream_sweep.err; echo "sweep rc=$?"; cat gpurun_out/${TAG}_stream_sweep.jsonl; tail -3 gpurun_out/${TAG}_stream_sweep.err

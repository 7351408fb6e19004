#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cpp_host.py tests/test_gpu_fullsize.py -x -q -k "readable or stream or cpp or committed_fixture or literal or config3 or concurrent" > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2j_tests.log
timeout 300 python tools/bench_stream_sweep.py > gpurun_out/r2j_stream_sweep.jsonl 2> gpurun_out/r2j_stream_sweep.err; echo "sweep rc=$?"; cat gpurun_out/r2j_stream_sweep.jsonl | cut -c1-600; tail -2 gpurun_out/r2j_stream_sweep.err
timeout 300 python tools/bench_stream_sweep.py > gpurun_out/r2j_stream_sweep2.jsonl 2>/dev/null; cat gpurun_out/r2j_stream_sweep2.jsonl | cut -c1-600
timeout 600 python bench.py --config 5 --steps 3 --warmup 2 > gpurun_out/r2j_bench_cfg5.jsonl 2> gpurun_out/r2j_bench_cfg5.err; tail -2 gpurun_out/r2j_bench_cfg5.err; cut -c1-400 gpurun_out/r2j_bench_cfg5.jsonl

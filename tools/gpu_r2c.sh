#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/mb_port > gpurun_out/r2c_mb_port.txt 2>&1; cat gpurun_out/r2c_mb_port.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tier_path or literal or fuzz_small or baseline_configs or full_1m or device_range or random_dictionaries" > gpurun_out/r2c_tests_a.log 2>&1; tail -5 gpurun_out/r2c_tests_a.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q --durations=12 > gpurun_out/r2c_tests_b.log 2>&1; tail -25 gpurun_out/r2c_tests_b.log
bash tools/gpu_exp.sh r2c

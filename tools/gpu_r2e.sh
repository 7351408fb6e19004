#!/bin/bash
mkdir -p gpurun_out
K="tier_path or literal or full_1m or device_range or baseline_configs or sel2 or compact or random_dictionaries"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "$K" > gpurun_out/r2e_tests_product.log 2>&1; tail -3 gpurun_out/r2e_tests_product.log
ACGPU_LIB=$PWD/ahocorasick_b200/variants/libacgpu_tma.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "$K" > gpurun_out/r2e_tests_tma.log 2>&1; tail -3 gpurun_out/r2e_tests_tma.log
VARIANTS="tma" LAUNCHES=1 bash tools/gpu_exp.sh r2e
FULLVAR=tma FULL="k_tier_mask" VARIANTS="" bash tools/gpu_exp.sh r2e_tma | tail -2

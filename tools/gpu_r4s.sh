#!/bin/bash
# compute-sanitizer over the kernels that are new in the last sessions: k_ww3_hits / k_ww3_emit, k_segments, chain-shard feeds
TAG=${1:-r4s}
mkdir -p gpurun_out
SEL='test_wholeword_custom_word_chars_config3_style or (test_keywords_longer_than_the_selection_kernels_hold and (2048 or 255)) or (not_closed_under_lowercase and 0]) or (test_readable_streaming_chain_blocks and 8192) or test_ww_hash_path_shapes'
timeout 1000 compute-sanitizer --tool memcheck --error-exitcode 99 --target-processes all python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${TAG}_memcheck.log | tail -3
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 99 --target-processes all python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_wholeword_custom_word_chars_config3_style or test_ww_hash_path_shapes" > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/${TAG}_racecheck.log | tail -3

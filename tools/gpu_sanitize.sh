#!/bin/bash
# compute-sanitizer (memcheck, then racecheck) over the small-input parity tests: every kernel family runs at least once.
# usage: tools/gpu_sanitize.sh <tag>
TAG=${1:-san}
mkdir -p gpurun_out
SEL='test_literal_cases or test_early_stop_quirks or test_wholeword_custom_word_chars_config3_style or test_committed_fixture_streams'
for TOOL in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 99 --target-processes all \
      python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/${TAG}_$TOOL.log 2>&1
  echo "$TOOL rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/${TAG}_$TOOL.log | tail -4
done

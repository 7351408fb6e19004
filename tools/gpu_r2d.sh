#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/mb_port > gpurun_out/r2d_mb_port.txt 2>&1; cat gpurun_out/r2d_mb_port.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "compact or tier_path or literal or full_1m or device_range" > gpurun_out/r2d_tests_a.log 2>&1; tail -15 gpurun_out/r2d_tests_a.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q --durations=12 > gpurun_out/r2d_tests_b.log 2>&1; tail -25 gpurun_out/r2d_tests_b.log
timeout 600 python bench.py --haystacks 1 --steps 3 --warmup 2 > gpurun_out/r2d_bench1.json 2> gpurun_out/r2d_bench1.err; tail -3 gpurun_out/r2d_bench1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2d_bench1.json').read())
print('value', d['value'], 'frac', d['roofline']['frac'], 'launch_ms', d['roofline']['launch_ms'])
print('e2e', json.dumps(d['e2e'])[:900])
print('parity', d.get('parity'), 'cpu', d.get('cpu_baseline'))
PY
LAUNCHES=1 FULL="k_tier_emit" bash tools/gpu_exp.sh r2d

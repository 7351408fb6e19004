#!/bin/bash
# Quick GPU check: parity tests (optional), short bench, per-kernel launch times, optional full ncu capture of one kernel.
# usage: tools/gpu_quick.sh <tag> [tests|notests] [kernel-regex-for-full-capture]
TAG=${1:-q}
mkdir -p gpurun_out
if [ "$2" == "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${TAG}_tests.log
  tail -5 gpurun_out/${TAG}_tests.log
fi
SHORT="python bench.py --haystacks 2 --chars 1000000000 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 600 $SHORT > gpurun_out/${TAG}_short.json 2> gpurun_out/${TAG}_short.err; echo "short rc=$?"; cat gpurun_out/${TAG}_short.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline'])"
tail -3 gpurun_out/${TAG}_short.err
NCU="python bench.py --haystacks 1 --chars 1000000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_tier|k_row|k_ac' --csv --log-file gpurun_out/${TAG}_launches.csv $NCU > gpurun_out/${TAG}_launches.log 2>&1
grep -E "k_tier|k_row|k_ac" gpurun_out/${TAG}_launches.csv | awk -F'","' '{print substr($5,1,60), $NF}' | tail -8
if [ -n "$3" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s 1 -c 1 -f -o gpurun_out/${TAG}_prof $NCU > gpurun_out/${TAG}_prof.log 2>&1
  ls -la gpurun_out/${TAG}_prof.ncu-rep
fi

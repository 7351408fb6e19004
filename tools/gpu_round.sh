#!/bin/bash
# One GPU session: parity tests, default bench, ncu launch list and one full capture of the tier kernel.
# usage: tools/gpu_round.sh <tag> [skip-tests]
TAG=${1:-x}
mkdir -p gpurun_out
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${TAG}_tests.log
  tail -3 gpurun_out/${TAG}_tests.log
fi
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json
SHORT="python bench.py --haystacks 1 --chars 200000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv $SHORT > gpurun_out/${TAG}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ac_tier -s 2 -c 1 -f -o gpurun_out/${TAG}_prof $SHORT > gpurun_out/${TAG}_prof.log 2>&1
ls -la gpurun_out | tail -8

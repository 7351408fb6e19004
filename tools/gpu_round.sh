#!/bin/bash
# One GPU session: parity tests, default bench (both arms), ncu launch list and full captures of the hot kernels.
# usage: tools/gpu_round.sh <tag> [skip-tests]
TAG=${1:-x}
mkdir -p gpurun_out
if [ "$2" != "skip-tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${TAG}_tests.log
  tail -3 gpurun_out/${TAG}_tests.log
fi
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
cat gpurun_out/${TAG}_bench_ref.json
NCU="python bench.py --haystacks 1 --chars 1000000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_tier|k_row|k_ac|k_fwd|k_sel' --csv --log-file gpurun_out/${TAG}_launches.csv $NCU > gpurun_out/${TAG}_launches.log 2>&1
for KR in k_tier_mask k_tier_emit; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KR -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_$KR $NCU > gpurun_out/${TAG}_prof_$KR.log 2>&1
done
ls -la gpurun_out | tail -8

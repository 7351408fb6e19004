#!/bin/bash
# k_tier_duo (generation 6: masks of slab i || records of slab i-1): parity with ACGPU_DUO=1, then configs[4] against the product path
mkdir -p gpurun_out
TAG=${1:-r5a}
[ -n "$SKIPTESTS" ] || ACGPU_DUO=1 ACGPU_DUO_MIN_ROWS=8192 timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -x -q -k "${KEXPR:-config4 or config1 or full_1m or tier_path or baseline_configs or large_haystack or ahocorasick_positions or cfg4 or prefix}" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
SHORT="python bench.py --haystacks 2 --chars 1000000000 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline"
run() {
  timeout 600 $SHORT > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err; rc=$?
  python -c "import sys,json; d=json.loads(open('gpurun_out/${TAG}_$1.json').read()); r=d['roofline']; print('$1 rc=$rc launch_ms %.3f frac %.3f matches %d clk %s parity %s' % (r['launch_ms'], r['frac'], d['matches_per_step'], d['clocks']['sm_mhz'], d.get('parity_checked')))" || tail -3 gpurun_out/${TAG}_$1.err
}
run product
for CFG in ${CFGS:-4:11 4:16 8:11 4:32}; do
  export ACGPU_DUO=1 ACGPU_DUO_SLABS=${CFG%%:*} ACGPU_DUO_EMIT_WARPS=${CFG#*:}
  run duo_s${ACGPU_DUO_SLABS}_e${ACGPU_DUO_EMIT_WARPS}
done
unset ACGPU_DUO

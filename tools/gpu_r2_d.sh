#!/bin/bash
mkdir -p gpurun_out
{
timeout 3000 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit=$?"
grep -E "passed|failed" gpurun_out/r2d_pytest.log | tail -3
grep -E "^FAILED|^ERROR|Error" gpurun_out/r2d_pytest.log | head -20
for tag in base nots im2old; do
  env=""
  [ $tag = nots ] && env="PESR_NO_TWO_STREAMS=1"
  [ $tag = im2old ] && env="PESR_IM2COL_OLD=1"
  env $env timeout 900 python bench.py --steps 50 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2d_bench_$tag.json 2> gpurun_out/r2d_bench_$tag.err; echo "bench $tag exit=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2d_bench_$tag.json'))
    print('$tag', 'ms/step', round(d['ms_per_step'],3), 'sustained', round(d['sustained']['ms_per_step'],3), 'launch', d['config']['launch'], d.get('graph_note'), 'clk', d['clocks']['sm_mhz'], d['sustained']['clocks']['sm_mhz'])
except Exception as e:
    print('$tag failed', e)
PY
  tail -3 gpurun_out/r2d_bench_$tag.err
done
} > gpurun_out/r2d.log 2>&1
tail -40 gpurun_out/r2d.log

#!/bin/bash
mkdir -p gpurun_out
{
for tag in s227 s200 s176 s176nots; do
  env=""
  [ $tag = s200 ] && env="PESR_CONV_SMEM_KB=200"
  [ $tag = s176 ] && env="PESR_CONV_SMEM_KB=176"
  [ $tag = s176nots ] && env="PESR_CONV_SMEM_KB=176 PESR_NO_TWO_STREAMS=1"
  env $env timeout 900 python bench.py --steps 50 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2e_bench_$tag.json 2> gpurun_out/r2e_bench_$tag.err; echo "bench $tag exit=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2e_bench_$tag.json'))
    print('$tag', 'ms/step', round(d['ms_per_step'],3), 'sustained', round(d['sustained']['ms_per_step'],3), 'launch', d['config']['launch'], d.get('graph_note'), 'clk', d['clocks']['sm_mhz'], d['sustained']['clocks']['sm_mhz'])
except Exception as e:
    print('$tag failed', e)
PY
  tail -3 gpurun_out/r2e_bench_$tag.err
done
} > gpurun_out/r2e.log 2>&1
tail -40 gpurun_out/r2e.log

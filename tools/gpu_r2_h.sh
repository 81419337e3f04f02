#!/bin/bash
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests/test_gan_gpu.py tests/test_generator_gpu.py tests/test_pinned_gradients_gpu.py tests/test_boundary_gpu.py tests/test_headline_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -4
for tag in lane nolane; do
  env=""
  [ $tag = nolane ] && env="PESR_NO_WGRAD_STREAM=1"
  for wl in gan pretrain; do
  env $env timeout 900 python bench.py --workload $wl --steps 50 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2h_bench_${tag}_$wl.json 2> gpurun_out/r2h_bench_${tag}_$wl.err; echo "bench $tag $wl exit=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2h_bench_${tag}_$wl.json'))
    print('$tag $wl', 'ms/step', round(d['ms_per_step'],3), 'sustained', round(d['sustained']['ms_per_step'],3), d['config']['launch'], d.get('graph_note'), 'clk', d['clocks']['sm_mhz'], 'host', round(d['host_enqueue_ms_per_step'],2))
except Exception as e:
    print('$tag $wl failed', e)
PY
  tail -2 gpurun_out/r2h_bench_${tag}_$wl.err
  done
done
PESR_NO_WGRAD_STREAM=0 timeout 600 python bench.py --steps 30 --warmup 3 --no-extras --no-cpu-baseline --no-graph | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('eager lane: ms/step', d['ms_per_step'], 'host enqueue', d['host_enqueue_ms_per_step'])"
} > gpurun_out/r2h.log 2>&1
tail -30 gpurun_out/r2h.log

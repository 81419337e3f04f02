#!/bin/bash
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests/test_netops_gpu.py tests/test_gan_gpu.py tests/test_conv_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
timeout 900 python bench.py --steps 50 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench exit=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2g_bench.json'))
print('ms/step', round(d['ms_per_step'],3), 'sustained', round(d['sustained']['ms_per_step'],3), d['config']['launch'], 'clk', d['clocks']['sm_mhz'])
PY
timeout 1200 python tools/torch_gpu_baseline.py 10 > gpurun_out/r2g_torch_gpu.json 2> gpurun_out/r2g_torch_gpu.err; echo "torch baseline exit=$?"
tail -3 gpurun_out/r2g_torch_gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2g_torch_gpu.json'))
for k in ('fp32_tf32','bf16_autocast_channels_last'): print(k, d[k])
for r in d['conv_table_fp16_channels_last']:
    print('%-36s cudnn f/d/w %7.1f %7.1f %7.1f us | ours %7.1f %7.1f %7.1f us' % (r['shape'], r['cudnn_us']['fprop'], r['cudnn_us']['dgrad'], r['cudnn_us']['wgrad'], r['pesr_b200_us']['fprop'], r['pesr_b200_us']['dgrad'], r['pesr_b200_us']['wgrad_incl_reduce']))
PY
} > gpurun_out/r2g.log 2>&1
tail -40 gpurun_out/r2g.log

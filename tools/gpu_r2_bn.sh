#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_netops_gpu.py tests/test_gan_gpu.py tests/test_pinned_gradients_gpu.py -q -m gpu -p no:cacheprovider -x -k "batchnorm or discriminator or gan_step_matches_pinned_oracle" 2>&1 | tail -3
timeout 300 python tools/perf_chain.py 2>&1 | grep -A8 "BatchNorm kernels"
timeout 300 python bench.py --steps 50 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('GAN', d['ms_per_step'], d['sustained']['ms_per_step'], d['clocks']); [print(' ', h['kernel'][:60], round(h['us'],1), round(h['frac'],2)) for h in d['roofline']['hbm'] if 'bn_' in h['kernel']]"
} > gpurun_out/r2bn.log 2>&1
cat gpurun_out/r2bn.log | cut -c1-300

#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3; echo "exit=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('GAN', d['value'], d['ms_per_step'], 'igemm', d['roofline']['achieved'], d['roofline']['share_of_step'], 'wgrad', d['roofline']['wgrad']['achieved'], d['roofline']['wgrad']['share_of_step'])
print('pretrain', d['pretrain_step']['value'], d['pretrain_step']['ms_per_step'], 'infer', d['inference_alpha1']['339x510'])
"; echo "bench gan exit=$?"
} > gpurun_out/bringup13.log 2>&1
tail -30 gpurun_out/bringup13.log

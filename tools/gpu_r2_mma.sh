#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_netops_gpu.py tests/test_generator_gpu.py tests/test_pinned_gradients_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
timeout 300 python tools/perf_chain.py 2>&1 | grep -A4 "trunk conv variants\|wgrad only"
timeout 300 python tools/phase_times.py gan 10 detail 2>&1 | grep "gan step\|64->64\|128->64\|64->128\|M64 \|N64 \|VGG forward\|VGG backward\|D forward\|D backward\|G forward\|G backward"
timeout 300 python bench.py --steps 50 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('GAN', d['ms_per_step'], d['sustained']['ms_per_step'], d['clocks'])"
timeout 300 python bench.py --workload pretrain --steps 50 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('pretrain', d['ms_per_step'], d['sustained']['ms_per_step'])"
} > gpurun_out/r2mma.log 2>&1
cat gpurun_out/r2mma.log | cut -c1-200

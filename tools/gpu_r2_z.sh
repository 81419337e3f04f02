#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_generator_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
PESR_B200_LIB=pesr_b200/libpesr_b200_debug.so timeout 200 python tools/perf_pack.py 2>&1 | tail -3
timeout 300 python bench.py --steps 50 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('GAN', d['ms_per_step'], d['sustained']['ms_per_step'], d['clocks'])"
} > gpurun_out/r2z.log 2>&1
cat gpurun_out/r2z.log | cut -c1-300

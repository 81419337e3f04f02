#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_netops_gpu.py tests/test_gan_gpu.py tests/test_pinned_gradients_gpu.py -q -m gpu -p no:cacheprovider -x -k "linear or discriminator or gan_step" -s 2>&1 | grep -v "^$" | tail -40
PESR_B200_LIB=pesr_b200/libpesr_b200_debug.so timeout 300 python tools/sm_hog_probe.py
timeout 300 python bench.py --steps 50 --warmup 3 --no-extras --no-cpu-baseline 2> gpurun_out/r2l_bench.err | tee gpurun_out/r2l_bench.json | cut -c1-400
PESR_FC1_WGRAD_SKINNY=1 timeout 300 python bench.py --steps 50 --warmup 3 --no-extras --no-cpu-baseline 2>> gpurun_out/r2l_bench.err | tee gpurun_out/r2l_bench_skinny.json | cut -c1-200
} > gpurun_out/r2l.log 2>&1
tail -60 gpurun_out/r2l.log | cut -c1-400

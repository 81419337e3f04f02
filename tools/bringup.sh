#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_netops_gpu.py tests/test_generator_gpu.py -x -q -m gpu 2>&1 | tail -4; echo "exit=$?"
timeout 300 python tools/perf_conv.py conv | grep wgrad; echo "exit=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline | cut -c1-330; echo "bench exit=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --workload gan --no-cpu-baseline | cut -c1-330; echo "bench gan exit=$?"
} > gpurun_out/bringup9.log 2>&1
tail -40 gpurun_out/bringup9.log

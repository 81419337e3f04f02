#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_netops_gpu.py -x -q -m gpu 2>&1 | tail -4; echo "exit=$?"
timeout 300 python tools/perf_conv.py pair; echo "exit=$?"
timeout 300 python tools/perf_conv.py conv | grep wgrad; echo "exit=$?"
} > gpurun_out/bringup7.log 2>&1
tail -40 gpurun_out/bringup7.log

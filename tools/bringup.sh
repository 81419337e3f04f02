#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_netops_gpu.py -x -q -m gpu 2>&1 | tail -15; echo "exit=$?"
timeout 300 python tools/perf_conv.py pair; echo "exit=$?"
} > gpurun_out/bringup4.log 2>&1
tail -60 gpurun_out/bringup4.log

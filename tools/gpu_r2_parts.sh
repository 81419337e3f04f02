#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_netops_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -4
timeout 120 python tools/perf_narrow.py 2>&1 | tail -26
} > gpurun_out/r2parts.log 2>&1
cat gpurun_out/r2parts.log | cut -c1-200

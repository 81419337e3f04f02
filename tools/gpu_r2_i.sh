#!/bin/bash
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests/test_cli_gpu.py tests/test_split_precision_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -30
} > gpurun_out/r2i.log 2>&1
tail -40 gpurun_out/r2i.log | cut -c1-300
